"""Summarise an ncu report (read here, without a GPU) into markdown for profiles/.
usage: python tools/ncu_summary.py <report.ncu-rep> [<launches.csv>] > profiles/<name>.md"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print("# ncu summary: %s\n" % rep.split("/")[-1])
    print("Captured with `ncu --set full --clock-control none --import-source on` under gpurun (cold-cache, serialised")
    print("replays: durations are NOT bench values; bench numbers come from CUDA events in bench.py).\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("## %s\n" % name[:160])
        print("| metric | value | unit |\n|---|---|---|")
        for i, h in enumerate(hdr):
            if h in KEYS or "tensor" in h and "pct_of_peak_sustained_active" in h and "cycles_active.avg" in h:
                print("| %s | %s | %s |" % (h, r[i], units[i]))
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("\nTop warp stall reasons (warps stalled per issue-active cycle): " +
              ", ".join("%s %.2f" % (n, v) for v, n in stalls[:7]) + "\n")
    if len(sys.argv) > 2:
        print("## Launch list of the bench command (`--metrics gpu__time_duration.sum`): share of the step\n")
        tot = {}
        with open(sys.argv[2]) as f:
            lines = [ln for ln in f if ln.startswith('"')]
        for row in csv.DictReader(lines):
            if row.get("Metric Name") != "gpu__time_duration.sum":
                continue
            k = row["Kernel Name"].split("(")[0][-70:]
            v = float(row["Metric Value"].replace(",", ""))
            u = row["Metric Unit"]
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
            c = tot.setdefault(k, [0, 0.0])
            c[0] += 1
            c[1] += v
        total = sum(v for _, v in tot.values())
        print("| kernel | launches | total ms | share |\n|---|---|---|---|")
        for k, (n, v) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:12]:
            print("| %s | %d | %.3f | %.1f %% |" % (k, n, v, 100 * v / total))


if __name__ == "__main__":
    main()
