#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for the Eigen dense-product hot path on B200.

Metric (BASELINE.json): dgemm TFLOP/s at N = 16384 (`C += A*B`, alpha = beta = 1, uniform[-1,1] inputs, C = 1 as in
bench/bench_gemm.cpp:130-133,211-214), 2*m*n*k flops per step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload dgemm16384|...]

* ours, N = 1: `value` = device-resident throughput (operands in HBM, one b200blas_gemm_dev call per step, CUDA
  events on the launching stream); `e2e` = the same product through the drop-in F77 entry point dgemm_ from ordinary
  PAGEABLE host matrices (what Eigen matrices are), host<->device copies inside the timed region; `e2e_pinned` = the same
  with the operands page-locked first; `roofline` = FP64 DMMA pipe (tensor bound); `cpu_baseline` = the reference's OpenMP
  gebp path (oracle/_ref) on this box's host cores on a bounded slab; `checked` = sampled rows of the timed result against
  the long-double oracle; `configs` = BASELINE.json's C1/C3/C4/C5 (incl. the nine op pairs of sgemm 8192^3).
* ours, N > 1 (torchrun, one rank per GPU; NCCL for the barriers and the max-over-ranks reduction): the partitioner
  lives BEHIND the C ABI (eigen_b200/csrc/multi.cu): rank 0 holds the operands (root-resident) and one dgemm_ /
  b200blas_gemm_dev call drives all N GPUs -- 2-D tiles of C, panels relayed by peer-to-peer copy-engine transfers, C
  sub-slabs returned to the root; strong scaling (total work fixed).  `presharded` = the same grid with panels and tiles
  already resident, one process per GPU (max over ranks).
* --workload <t><routine><n> with routine in syrk/trsm/trmm/symm/syr2k/potrf/getrf (SURVEY 8 f rows, single GPU,
  diagnostic -- never the driver's default): one F77 call per step on device pointers (`value`) and on pinned host
  operands (`e2e`); `cpu_baseline` / `--impl reference` time the reference's own blas/ or lapack/ routine (oracle/_ref).
* --impl reference: the reference's own CPU implementation of the path (oracle/_ref/libeigen_gebp_omp.so, Eigen's
  expression API + OpenMP gebp) on the host cores, each step a bounded column slab of the workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (type, m, n, k, alpha, beta, lda/ldb/ldc override or None)
    "dgemm16384": ("d", 16384, 16384, 16384, 1.0, 1.0),
    "dgemm8192": ("d", 8192, 8192, 8192, 1.0, 1.0),
    "dgemm2048": ("d", 2048, 2048, 2048, 1.0, 1.0),
    "sgemm8192": ("s", 8192, 8192, 8192, 1.0, 1.0),
    "sgemm16384": ("s", 16384, 16384, 16384, 1.0, 1.0),
    "zgemm4096": ("z", 4096, 4096, 4096, 1.0, 1.0),
    "cgemm4096": ("c", 4096, 4096, 4096, 1.0, 1.0),
    "dgemm_rankk": ("d", 16384, 16384, 256, -1.0, 1.0),
}
FLOP_FACTOR = {"s": 2.0, "d": 2.0, "c": 8.0, "z": 8.0}
ESIZE = {"s": 4, "d": 8, "c": 8, "z": 16}
DTYPE_NAME = {"s": "f32 (3xTF32 split, fp32 accumulate)", "d": "f64", "c": "c64 (3xTF32 split)", "z": "c128 (f64)"}


def torch_dtype(t):
    import torch
    return {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128}[t]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0, ngpus=1):
        self.index = index          # None: sample GPUs 0 .. ngpus-1 (rank 0 drives every device)
        self.ngpus = ngpus
        self.proc = None
        self.lines = []
        self.mark_at = 0

    def mark(self):
        """Call at the start of the timed region: only samples taken after this point are reported (the sampler is
        started before the warm-up because nvidia-smi needs ~0.3 s to deliver its first line)."""
        self.mark_at = len(self.lines)

    def start(self):
        try:
            sel = ["-i", str(self.index)] if self.index is not None else ["-i", ",".join(str(i) for i in range(self.ngpus))]
            self.proc = subprocess.Popen(["nvidia-smi"] + sel + ["--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        timed = self.lines[self.mark_at:]
        window = "timed region"
        if not timed:   # region shorter than one sampling period: fall back to the warm-up samples (same load)
            timed, window = self.lines, "warm-up + timed region"
        for ln in timed:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm_sorted = sorted(sm)
        # "under load": the upper half of the samples belongs to the timed region when the GPU was busy
        busy = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
        busy.sort()
        return {"sm_mhz": busy[len(busy) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "power_w_max": max(power), "samples": len(sm), "sm_mhz_min": sm_sorted[0], "window": window}


def cpu_slab_columns(t, m, n, k, cores):
    """ONE rule for both CPU legs (`cpu_baseline` and `--impl reference`): at least 256 columns of C per core -- so that
    parallelize_gemm (Parallelizer.h:108-151) gives every thread a full-width share, the regime the 16384-column product
    runs in -- capped by the width of the workload."""
    return max(256, min(n, 256 * max(1, cores)))


def cpu_reference_leg(t, m, n, k, alpha, beta, steps, warmup, slab_cols):
    """Times the reference's OpenMP gebp path (Eigen expression API -> parallelize_gemm -> gebp_kernel) on a
    column slab of the workload: C[:, :slab] += A * B[:, :slab].  Returns (TFLOP/s, cores, sample, kind, ms/step)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api as oa
    cores = os.cpu_count() or 1
    ns = min(n, slab_cols)
    rng = np.random.default_rng(42)
    dt = oa.NP_DTYPE[t]
    A = host_uniform(np, rng, t, m, k)
    B = host_uniform(np, rng, t, k, ns)
    Cm = np.ones((m, ns), dtype=dt, order="F")
    al = np.array([alpha], dtype=dt)
    be = np.array([beta], dtype=dt)
    if oa.have_ref():
        kind = "reference"
        S = oa.ref_shim()
        fn = getattr(S, "ref_eigen_gemm_" + t)

        def step():
            fn(b"N", b"N", m, ns, k, oa._ptr(al), oa._ptr(A), m, oa._ptr(B), k, oa._ptr(be), oa._ptr(Cm), m, cores)
    else:
        kind = "port"
        P = oa.port()

        def step():
            P.oracle_gemm_omp(oa.TYPES[t], b"N", b"N", m, ns, k, oa._ptr(al), oa._ptr(A), m, oa._ptr(B), k, oa._ptr(be),
                              oa._ptr(Cm), m, cores)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt_s = (time.perf_counter() - t0) / max(steps, 1)
    flops = FLOP_FACTOR[t] * m * ns * k
    sample = ("%sgemm %dx%dx%d column slab (%d of %d columns = 256 per core) of the workload, C+=A*B, OMP threads=%d"
              % (t, m, ns, k, ns, n, cores))
    return flops / dt_s / 1e12, cores, sample, kind, dt_s * 1e3


def host_uniform(np, rng, t, rows, cols):
    """Ordinary (pageable, malloc'ed) column-major rows x cols matrix, uniform[-1,1] like Eigen's setRandom
    (bench/bench_gemm.cpp:211-214, MathFunctions.h:628-637)."""
    dt = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}[t]
    x = np.empty((rows, cols), dtype=dt, order="F")
    flat = x.reshape(-1, order="F")
    blk = 1 << 22
    for i in range(0, flat.size, blk):
        j = min(flat.size, i + blk)
        if t in "cz":
            flat[i:j] = (rng.uniform(-1, 1, j - i) + 1j * rng.uniform(-1, 1, j - i)).astype(dt)
        else:
            flat[i:j] = rng.uniform(-1, 1, j - i).astype(dt)
    return x


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    t, m, n, k, alpha, beta = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    slab = cpu_slab_columns(t, m, n, k, cores)
    # the driver's K + W steps must end within a few minutes: bound the step count, never the sample
    est_step_s = FLOP_FACTOR[t] * m * slab * k / (cores * 0.035e12)
    steps = max(1, min(args.steps, int(240.0 / max(est_step_s, 1e-3)) - args.warmup))
    warmup = args.warmup
    tf, cores, sample, kind, ms = cpu_reference_leg(t, m, n, k, alpha, beta, steps, warmup, slab)
    line = {
        "impl": "reference", "metric": metric_name(args.workload), "value": tf, "unit": "TFLOP/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": DTYPE_NAME[t].split()[0],
        "data": "synthetic", "config": workload_config(args.workload, args.gpus),
        "cpu_baseline": {"value": tf, "unit": "TFLOP/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": tf, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "steps_requested": args.steps,
    }
    print(json.dumps(line))
    return 0


def metric_name(workload):
    t, m, n, k = WORKLOADS[workload][:4]
    if workload == "dgemm16384":
        return "dgemm TFLOP/s at N=16384"
    return "%sgemm TFLOP/s at %dx%dx%d" % (t, m, n, k)


def grid_label(t, m, n, k, gpus):
    if gpus == 1:
        return "1 GPU"
    try:
        import eigen_b200
        info, _ = eigen_b200.multi_plan(t, "N", "N", m, n, k, 1.0, 1.0, gpus)
        grid = "%dx%d" % (info.pr, info.pc)
    except Exception:
        grid = "?"
    return ("%s grid of C tiles over %d GPUs, k unsplit; ONE process drives all devices through the C ABI (B200BLAS_NGPUS=%d): "
            "panel k-chunks fetched once per grid row / column and relayed by peer-to-peer copy-engine transfers, C sub-slabs "
            "returned to the root" % (grid, gpus, gpus))


def workload_config(workload, gpus):
    t, m, n, k, alpha, beta = WORKLOADS[workload]
    return {"workload": "%s: %sgemm C(%dx%d) = %g*A(%dx%d)*B(%dx%d) + %g*C, column-major, ld=rows, uniform[-1,1], C=1"
            % (workload, t, m, n, alpha, m, k, k, n, beta),
            "l2": ("inputs (%.2f GB) are larger than L2 (126 MB); no explicit flush" if ESIZE[t] * (m * k + k * n + m * n) > 2.5e8
                   else "inputs (%.2f GB) fit in L2: parity/diagnostic workload, not a bench line") % (ESIZE[t] * (m * k + k * n + m * n) / 1e9),
            "parallelism": grid_label(t, m, n, k, gpus)}


def sampled_row_check(t, m, n, k, alpha, beta, hA, hB, C0_value, got_rows, rows):
    """Sampled rows of a finished product against the long-double oracle (oracle/hp_ref.c; test infrastructure, outside
    every timed region).  hA: m x k, hB: k x n numpy (column-major); C started as the constant C0_value."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api as oa
    C0 = np.full((m, n), C0_value, dtype=oa.NP_DTYPE[t], order="F")
    ref, g = oa.hp_gemm(t, "N", "N", m, n, k, alpha, hA, m, hB, k, beta, C0, m, rows=np.asarray(rows, dtype=np.int32))
    ratio = float((np.abs(got_rows - ref) / (oa.EPS[t] * g)).max())
    fro = float(np.linalg.norm(got_rows - ref) / np.linalg.norm(ref))
    return {"rows": [int(r) for r in rows], "gauge_ratio": ratio, "rel_fro": fro, "bound_k_eps": k * oa.EPS[t],
            "ok": bool(ratio < 16.0 and fro <= k * oa.EPS[t]),
            "against": "long-double oracle (oracle/hp_ref.c), netlib gauge |c-ref|/(eps*G) < 16 (dblat3.f:2587-2596)"}


def max_over_ranks(value, device="cuda", group=None):
    """Contract: every multi-GPU time is the MAX over ranks (rank 0 drives all devices through the C ABI, the other ranks
    report 0 for that leg; in the pre-sharded leg every rank times its own tile)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    x = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(x, op=dist.ReduceOp.MAX, group=group)
    return float(x.item())


def time_device_calls(torch, call, steps, warmup, stream):
    for _ in range(warmup):
        call()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record(stream)
    for i in range(steps):
        call()
        ev[i + 1].record(stream)
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]


def extra_configs(torch, eigen_b200, peaks):
    """BASELINE.json configs C1, C3, C4, C5 (device-resident, one GPU) carried in the driver-run line: value, fraction
    of the measured pipe peak, clocks.  sgemm 8192^3 runs all nine op pairs (blas/testing/sblat3.dat grid)."""
    out = []
    gen = torch.Generator(device="cuda")
    gen.manual_seed(7)

    def urand(t, rows, cols):
        dt = torch_dtype(t)
        if t in "cz":
            re = torch.rand(cols, rows, 2, dtype=torch.float64 if t == "z" else torch.float32, device="cuda", generator=gen) * 2 - 1
            return torch.view_as_complex(re)
        return torch.rand(cols, rows, dtype=dt, device="cuda", generator=gen) * 2 - 1

    def one(name, t, ta, tb, m, n, k, alpha, beta, A, lda, B, ldb, Cd, ldc, note=None):
        stream = torch.cuda.current_stream()
        flops = FLOP_FACTOR[t] * m * n * k

        def call():
            r = eigen_b200.gemm_dev(t, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, Cd, ldc, stream=stream.cuda_stream)
            assert r == 0, eigen_b200.last_error()

        call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); call(); e1.record(stream)
        torch.cuda.synchronize()
        est = max(e0.elapsed_time(e1), 1e-3)
        steps = int(max(5, min(400, 600.0 / est)))     # ~0.6 s per config so that the clock sampler sees it
        clocks = ClockSampler(torch.cuda.current_device())
        clocks.start()
        time.sleep(0.35)                               # nvidia-smi needs ~0.3 s for its first line
        clocks.mark()
        per = time_device_calls(torch, call, steps, 3, stream)
        clk = clocks.stop()
        ms = sum(per) / len(per)
        variant = eigen_b200.last_variant()
        issued = flops * (3.0 if t in "sc" and "tf32" in variant else 1.0)
        peak = peaks["dmma" if t in "dz" else "tf32"]
        row = {"config": name, "op": ta + tb, "value": flops / (ms * 1e-3) / 1e12, "unit": "TFLOP/s", "ms_per_step": ms,
               "ms_best": min(per), "steps": steps, "kernel": variant,
               "frac": issued / (ms * 1e-3) / 1e12 / peak if peak and peak > 0 else None,
               "frac_of": "%s pipe peak measured in this run (%.1f TFLOP/s)%s" % ("FP64 DMMA" if t in "dz" else "TF32 tcgen05", peak or 0.0,
                                                                                   "; 3 issued tensor flops per algorithmic flop" if issued != flops else ""),
               "clocks": clk}
        if note:
            row["note"] = note
        out.append(row)

    # C3: sgemm 8192^3, all nine op pairs
    n = 8192
    A, B, Cd = urand("s", n, n), urand("s", n, n), torch.ones(n, n, dtype=torch.float32, device="cuda")
    for ta in "NTC":
        for tb in "NTC":
            one("C3 sgemm 8192^3", "s", ta, tb, n, n, n, 1.0, 1.0, A, n, B, n, Cd, n)
    del A, B, Cd
    # C4: complex 4096^3
    n = 4096
    for t in "zc":
        A, B, Cd = urand(t, n, n), urand(t, n, n), urand(t, n, n)
        one("C4 %sgemm 4096^3" % t, t, "N", "N", n, n, n, 1.0, 1.0, A, n, B, n, Cd, n)
        one("C4 %sgemm 4096^3" % t, t, "C", "N", n, n, n, 0.7 - 0.9j, 1.3 - 1.1j, A, n, B, n, Cd, n, note="alpha/beta of zblat3.dat:12-14")
        del A, B, Cd
    # C5: rank-k trailing update on sub-blocks of ONE 16640^2 matrix (PartialPivLU.h:492), alpha = -1
    N, bs = 16640, 256
    M = urand("d", N, N)
    mm = N - bs
    one("C5 dgemm rank-k 16384x16384x256", "d", "N", "N", mm, mm, bs, -1.0, 1.0, M[0:bs, bs:], N, M[bs:, 0:bs], N, M[bs:, bs:], N,
        note="A22 -= A21*A12, lda = ldb = ldc = 16640")
    del M
    # C1 on the GPU (the CPU figure for this config is the reference arm's business)
    n = 2048
    A, B, Cd = urand("d", n, n), urand("d", n, n), torch.ones(n, n, dtype=torch.float64, device="cuda")
    one("C1 dgemm 2048^3", "d", "N", "N", n, n, n, 1.0, 1.0, A, n, B, n, Cd, n)
    del A, B, Cd
    torch.cuda.empty_cache()
    return out


def f77_gemm_call(L, t, m, n, k, alpha, beta, hA, hB, hC):
    import numpy as np
    npdt = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}[t]
    fn = getattr(L, t + "gemm_")
    al = np.array([alpha], dtype=npdt)
    be = np.array([beta], dtype=npdt)
    ints = [C.c_int(x) for x in (m, n, k, m, k, m)]

    def call():
        r = fn(b"N", b"N", C.byref(ints[0]), C.byref(ints[1]), C.byref(ints[2]), al.ctypes.data_as(C.c_void_p),
               C.c_void_p(hA.ctypes.data), C.byref(ints[3]), C.c_void_p(hB.ctypes.data), C.byref(ints[4]),
               be.ctypes.data_as(C.c_void_p), C.c_void_p(hC.ctypes.data), C.byref(ints[5]))
        assert r == 0
    call.keep = (al, be, ints)
    return call


def e2e_legs(L, torch, eigen_b200, t, m, n, k, alpha, beta, steps, ngpus):
    """The same product through the drop-in F77 entry point from HOST matrices, host<->device copies inside the timed
    region.  `e2e`: ordinary pageable memory (what Eigen matrices and bench_gemm.cpp:211-214 allocate); `e2e_pinned`: the
    same buffers page-locked with b200blas_host_register first.  Returns (e2e, e2e_pinned, hA, hB, last C) ."""
    import numpy as np
    rng = np.random.default_rng(42)
    hA, hB = host_uniform(np, rng, t, m, k), host_uniform(np, rng, t, k, n)
    npdt = hA.dtype
    hC = np.ones((m, n), dtype=npdt, order="F")
    flops = FLOP_FACTOR[t] * m * n * k
    call = f77_gemm_call(L, t, m, n, k, alpha, beta, hA, hB, hC)
    api = "%sgemm_ (F77 C ABI, include/b200blas.h)%s" % (t, "" if ngpus == 1 else ", B200BLAS_NGPUS=%d" % ngpus)
    res = []
    for pinned in (False, True):
        if pinned:
            for x in (hA, hB, hC):
                assert L.b200blas_host_register(C.c_void_p(x.ctypes.data), x.nbytes) == 0
        try:
            hC.fill(1)
            call()                      # warm-up: allocates the staging buffers / device images
            hC.fill(1)
            call()
            torch.cuda.synchronize()
            hC.fill(1)
            t0 = time.perf_counter()
            for _ in range(steps):
                call()
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) / steps * 1e3
            h2d, d2h = C.c_uint64(), C.c_uint64()
            L.b200blas_last_transfer(C.byref(h2d), C.byref(d2h))
            res.append({"value": flops / (ms * 1e-3) / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": h2d.value,
                        "d2h_bytes_per_step": d2h.value, "ms_per_step": ms, "steps": steps,
                        "api": api + (" on page-locked host operands (b200blas_host_register)" if pinned else
                                      " on pageable host operands (numpy/malloc, as bench_gemm.cpp:211-214)")})
        finally:
            if pinned:
                for x in (hA, hB, hC):
                    L.b200blas_host_unregister(C.c_void_p(x.ctypes.data))
    return res[0], res[1], hA, hB, hC, steps


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import eigen_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # ranks > 0 must wait on the HOST while rank 0 drives their GPUs: an NCCL barrier would spin in a kernel there
        host_group = dist.new_group(backend="gloo")
    L = eigen_b200.require_device()
    t, m, n, k, alpha, beta = WORKLOADS[args.workload]
    dt = torch_dtype(t)
    flops = FLOP_FACTOR[t] * m * n * k
    gen = torch.Generator(device="cuda")
    gen.manual_seed(42 + rank)

    def urand(rows, cols):
        # column-major rows x cols == row-major (cols, rows)
        if t in "cz":
            re = torch.rand(cols, rows, 2, dtype=torch.float64 if t == "z" else torch.float32, device="cuda", generator=gen) * 2 - 1
            return torch.view_as_complex(re)
        return torch.rand(cols, rows, dtype=dt, device="cuda", generator=gen) * 2 - 1

    def device_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def host_barrier():
        if world > 1:
            dist.barrier(group=host_group)

    result = {}
    check_rows = [0, m // 2 - 1, m // 2, m - 1]
    pipe = 0 if t in "dz" else 3
    stream = torch.cuda.current_stream()
    ms_per_step, n_launch, value, variant = None, 0, None, ""
    # ================================ the product, operands resident in HBM =========================================
    if rank == 0:
        if world > 1:
            assert eigen_b200.set_devices(world) == world, "rank 0 must see all %d GPUs" % world
        A, B = urand(m, k), urand(k, n)
        Cd = torch.ones(n, m, dtype=dt, device="cuda")

        def step():
            r = eigen_b200.gemm_dev(t, "N", "N", m, n, k, alpha, A, m, B, k, beta, Cd, m, stream=stream.cuda_stream)
            assert r == 0, eigen_b200.last_error()
    clocks = ClockSampler(None if world > 1 else local_rank, world)
    if rank == 0:
        clocks.start()
    if rank == 0:
        for _ in range(args.warmup):
            step()
    device_barrier()
    clocks.mark()
    if rank == 0:
        launches0 = eigen_b200.kernel_launches()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        ev[0].record(stream)
        for i in range(args.steps):
            step()
            ev[i + 1].record(stream)
        torch.cuda.synchronize()
        n_launch = eigen_b200.kernel_launches() - launches0
        per = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
        total_ms = ev[0].elapsed_time(ev[-1])
        variant = eigen_b200.last_variant()
    else:
        total_ms, per = 0.0, [0.0]
    host_barrier()
    if rank == 0:
        result["clocks"] = clocks.stop()
    device_barrier()
    total_ms = max_over_ranks(total_ms)     # rank 0 drives every device; the others report 0
    ms_per_step = total_ms / args.steps
    value = flops / (ms_per_step * 1e-3) / 1e12
    if rank == 0:
        # ---- correctness of what was just timed: C = 1 + (warmup + steps) * A*B, sampled rows vs the long-double oracle
        reps = args.warmup + args.steps
        hAr = np.asfortranarray(A.cpu().numpy().T)
        hBr = np.asfortranarray(B.cpu().numpy().T)
        got = Cd[:, check_rows].cpu().numpy().T
        try:
            result["checked"] = sampled_row_check(t, m, n, k, alpha * reps, 1.0, hAr, hBr, 1.0, got, check_rows) if beta == 1.0 else None
            if result["checked"]:
                result["checked"]["what"] = "device-resident result after %d accumulating steps (C = 1 + %d*A*B)" % (reps, reps)
        except Exception as e:
            result["checked"] = {"ok": False, "error": repr(e)[:200]}
        del hAr, hBr
        # ---- roofline of the dominant kernel ---------------------------------------------------------------
        peak = eigen_b200.pipe_peak(pipe, 1500)
        peak_src = ("b200blas_pipe_peak(%s) measured in this run on this GPU; MEASURED_PEAKS.json has no %s figure"
                    % ("FP64 DMMA mma.sync.m8n8k4" if pipe == 0 else "TF32 tcgen05.mma", "FP64" if pipe == 0 else "TF32"))
        if peak <= 0 and pipe == 3:
            peak = 1100.0
            peak_src = "nominal dense TF32 1.1 PFLOP/s (fallback; microbenchmark unavailable)"
        issued = flops * (3.0 if t in "sc" and "tf32" in variant else 1.0)
        avg_ms = sum(per) / len(per)
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic_%s.json" % args.workload)
        if os.path.exists(tp) and world == 1:
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_step")
            except Exception:
                traffic = None
        per_step = max(1, n_launch // max(args.steps, 1))
        result["roofline"] = {
            "bound": "tensor", "achieved": issued / (avg_ms * 1e-3) / 1e12, "peak": peak * world, "unit": "TFLOP/s",
            "frac": issued / (avg_ms * 1e-3) / 1e12 / (peak * world) if peak > 0 else None, "traffic": traffic,
            "kernel": variant, "step_ms_avg": avg_ms, "step_ms_best": min(per), "launches_per_step": per_step,
            "note": "one step = one product = %d kernel launch(es) (k-slices / k-chunk groups / pack + product%s); achieved, traffic "
                    "and the algorithmic figures are per step%s" % (per_step, "; all devices" if world > 1 else "",
                                                                     "; peak = %d x the single-GPU pipe peak" % world if world > 1 else ""),
            "algorithmic_flops_per_step": flops, "tensor_pipe_flops_per_step": issued,
            "algorithmic_bytes_per_step": ESIZE[t] * (m * k + k * n + 2 * m * n), "peak_source": peak_src,
        }
        del A, B, Cd
        torch.cuda.empty_cache()
    # ================================ pre-sharded residency (N > 1): every rank owns its panels and its tile ============
    if world > 1:
        info, _ = eigen_b200.multi_plan(t, "N", "N", m, n, k, alpha, beta, world)
        i, j = divmod(rank, info.pc)
        mi, nj = info.row_cut[i + 1] - info.row_cut[i], info.col_cut[j + 1] - info.col_cut[j]
        if rank == 0:
            eigen_b200.set_devices(1)       # the local product below must stay on this rank's GPU
        Ai, Bj = urand(mi, k), urand(k, nj)
        Cij = torch.ones(nj, mi, dtype=dt, device="cuda")

        def local_step():
            r = eigen_b200.gemm_dev(t, "N", "N", mi, nj, k, alpha, Ai, mi, Bj, k, beta, Cij, mi, stream=stream.cuda_stream)
            assert r == 0, eigen_b200.last_error()
        for _ in range(args.warmup):
            local_step()
        device_barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            local_step()
        e1.record(stream)
        device_barrier()
        pms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
        result["presharded"] = {"value": flops / (pms * 1e-3) / 1e12, "unit": "TFLOP/s", "ms_per_step": pms,
                                "residency": "A_i, B_j, C_ij already resident on GPU (i,j) of the %dx%d grid, one process per GPU, no "
                                             "transfer in the timed region (max over ranks)" % (info.pr, info.pc)}
        del Ai, Bj, Cij
        torch.cuda.empty_cache()
        if rank == 0:
            eigen_b200.set_devices(world)
    # ================================ e2e: the drop-in F77 entry point with HOST operands ================================
    device_barrier()
    if rank == 0:
        try:
            e2e_steps = max(1, min(args.steps, 3))
            e2e, e2e_pinned, hA, hB, hC, done_steps = e2e_legs(L, torch, eigen_b200, t, m, n, k, alpha, beta, e2e_steps, world)
            result["e2e"], result["e2e_pinned"] = e2e, e2e_pinned
            if beta == 1.0:
                chk = sampled_row_check(t, m, n, k, alpha * done_steps, 1.0, hA, hB, 1.0, hC[check_rows], check_rows)
                chk["what"] = "host result of the pinned e2e leg after %d accumulating %sgemm_ calls" % (done_steps, t)
                result["checked_e2e"] = chk
            del hA, hB, hC
        except Exception as e:  # never lose the device-resident number to a host-memory problem
            result["e2e"] = {"value": None, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": repr(e)[:300]}
    host_barrier()
    if rank == 0 and world == 1:
        # ---- CPU baseline: the reference's OpenMP gebp on this box's cores, same slab rule as --impl reference -------
        try:
            slab = cpu_slab_columns(t, m, n, k, os.cpu_count() or 1)
            tf, cores, sample, kind, _ = cpu_reference_leg(t, m, n, k, alpha, beta, 2, 1, slab)
            result["cpu_baseline"] = {"value": tf, "unit": "TFLOP/s", "cores": cores, "kind": kind, "sample": sample}
        except Exception as e:  # the baseline is reported, never required for the GPU number
            result["cpu_baseline"] = {"value": None, "unit": "TFLOP/s", "cores": os.cpu_count(), "kind": "port",
                                      "sample": "failed: %r" % (e,)}
        # ---- the other BASELINE.json configs, device-resident, in the same driver-run line ---------------------------
        if args.workload == "dgemm16384" and not args.no_configs:
            try:
                peaks = {"dmma": result["roofline"]["peak"] if pipe == 0 else eigen_b200.pipe_peak(0, 800),
                         "tf32": result["roofline"]["peak"] if pipe == 3 else max(eigen_b200.pipe_peak(3, 800), eigen_b200.pipe_peak(3, 800))}
                result["configs"] = extra_configs(torch, eigen_b200, peaks)
            except Exception as e:
                result["configs"] = [{"error": repr(e)[:300]}]
    if world > 1:
        result["residency"] = ("value / e2e: A, B, C resident on GPU 0 (root-resident) resp. in host memory; distribution and the return of C "
                               "inside the timed region.  presharded: panels and tiles already on their GPUs")
    if rank == 0:
        line = {
            "metric": metric_name(args.workload), "value": value, "unit": "TFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": DTYPE_NAME[t].split()[0], "data": "synthetic",
            "config": workload_config(args.workload, world), "gpu_launches": int(n_launch), "kernel": variant,
        }
        line.update(result)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ---- widened rows: ?syrk_ ?trsm_ ?trmm_ ?symm_ ?syr2k_ ?potrf_ ?getrf_ ------------------------------
# Same line shape as the headline workloads; one F77 call per step.  These are diagnostic workloads (single GPU): the
# driver's default run never selects them.  Measured lines live in profiles/bench_r01/level3_*.jsonl.
LEVEL3_ROUTINES = ("syrk", "trsm", "trmm", "symm", "syr2k", "potrf", "getrf")
LEVEL3_WORKLOADS = {"%s%s%d" % (t, r, n): (t, r, n) for t in "sd" for r in LEVEL3_ROUTINES for n in (2048, 8192, 16384)}


def level3_flops(r, n):
    return {"syrk": n ** 3 * 1.0, "trsm": n ** 3 * 1.0, "trmm": n ** 3 * 1.0, "symm": 2.0 * n ** 3, "syr2k": 2.0 * n ** 3,
            "potrf": n ** 3 / 3.0, "getrf": 2.0 * n ** 3 / 3.0}[r]


def level3_call(lib, t, r, n, pa, pb, pc, ipiv):
    """The F77 call of one step (side / uplo / trans = L / L / N, alpha = 1 (syrk: -1), beta = 0 (syrk: 1))."""
    rt = C.c_float if t == "s" else C.c_double
    one, zero, mone = rt(1.0), rt(0.0), rt(-1.0)
    nn = C.c_int(n)
    info = C.c_int(0)
    bn = C.byref(nn)
    f = getattr(lib, t + r + "_")
    if r == "syrk":
        return lambda: f(b"L", b"N", bn, bn, C.byref(mone), pa, bn, C.byref(one), pc, bn)
    if r in ("trsm", "trmm"):
        return lambda: f(b"L", b"L", b"N", b"N", bn, bn, C.byref(one), pa, bn, pb, bn)
    if r in ("symm", "syr2k"):
        c1, c2 = (b"L", b"L") if r == "symm" else (b"L", b"N")
        return lambda: f(c1, c2, bn, bn, C.byref(one), pa, bn, pb, bn, C.byref(zero), pc, bn)
    if r == "potrf":
        return lambda: f(b"L", bn, pa, bn, C.byref(info))
    return lambda: f(bn, bn, pa, bn, ipiv.ctypes.data_as(C.POINTER(C.c_int)), C.byref(info))


def level3_cpu_reference(t, r, cn):
    """The reference's own routine (oracle/_ref; its blas/ and lapack/ are single-threaded) at a bounded order cn."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api as oa
    npdt = np.float32 if t == "s" else np.float64
    rng = np.random.default_rng(7)
    m = rng.uniform(-1, 1, (cn, cn))
    if r == "potrf":
        a = np.asfortranarray((m @ m.T / cn + np.eye(cn)).astype(npdt))
    elif r in ("trsm", "trmm"):
        a = np.asfortranarray((m * (2.0 / cn) + 1.5 * np.eye(cn)).astype(npdt))
    else:
        a = np.asfortranarray(m.astype(npdt))
    b = np.asfortranarray(rng.uniform(-1, 1, (cn, cn)).astype(npdt))
    c = np.ones((cn, cn), dtype=npdt, order="F")
    cpiv = np.zeros(cn, dtype=np.int32)
    lib = oa.ref_lapack() if r in ("potrf", "getrf") else oa.ref_blas()
    call = level3_call(lib, t, r, cn, oa._ptr(a), oa._ptr(b), oa._ptr(c), cpiv)
    t0 = time.perf_counter()
    call()
    dt_s = time.perf_counter() - t0
    return {"value": level3_flops(r, cn) / dt_s / 1e12, "unit": "TFLOP/s", "cores": 1, "kind": "reference",
            "sample": "%s%s_ of oracle/_ref (the reference's blas/ and lapack/ are single-threaded) at n=%d" % (t, r, cn)}, dt_s * 1e3


def run_level3(args):
    t, r, n = LEVEL3_WORKLOADS[args.workload]
    metric = "%s%s TFLOP/s at n=%d" % (t, r, n)
    config = {"workload": "%s: %s%s_ order %d, side/uplo/trans = L/L/N, uniform[-1,1] operands (potrf: M M^T / n + I; trsm/trmm: "
                          "unit-scale triangle), column-major, ld = n" % (args.workload, t, r, n),
              "l2": "operands larger than L2 for n >= 8192; no explicit flush", "parallelism": "1 GPU"}
    dtype = "f64" if t == "d" else "f32"
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return 0
        cpu, ms = level3_cpu_reference(t, r, 2048)
        print(json.dumps({"impl": "reference", "metric": metric, "value": cpu["value"], "unit": "TFLOP/s", "n_gpus": args.gpus,
                          "steps": 1, "warmup": 0, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": dtype, "data": "synthetic", "config": config, "cpu_baseline": cpu,
                          "e2e": {"value": cpu["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return 0
    if int(os.environ.get("WORLD_SIZE", "1")) != 1:
        if int(os.environ.get("RANK", "0")) == 0:
            print(json.dumps({"metric": metric, "unavailable": "the level-3 / LAPACK workloads are single-GPU (replicas only)"}))
        return 0
    import numpy as np
    import torch
    import eigen_b200
    L = eigen_b200.require_device()
    dt = torch.float32 if t == "s" else torch.float64
    eigen_b200.pipe_peak(0, 1500)   # ramp the clocks before the denominators are measured
    pipe = 0 if t == "d" else 3
    peak = max(eigen_b200.pipe_peak(pipe, 800), eigen_b200.pipe_peak(pipe, 800))
    g = torch.Generator(device="cuda").manual_seed(7)
    M = torch.rand(n, n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    if r == "potrf":
        src = (M @ M.T / n + torch.eye(n, dtype=torch.float64, device="cuda")).to(dt)
    elif r in ("trsm", "trmm"):
        src = (M * (2.0 / n) + torch.eye(n, dtype=torch.float64, device="cuda") * 1.5).to(dt)
    else:
        src = M.to(dt)
    Bd = (torch.rand(n, n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1).to(dt)
    Cd = torch.ones(n, n, dtype=dt, device="cuda")
    del M
    ipiv = np.zeros(n, dtype=np.int32)
    inplace_a = r in ("potrf", "getrf")
    A = src.clone()
    call = level3_call(L, t, r, n, C.c_void_p(A.data_ptr()), C.c_void_p(Bd.data_ptr()), C.c_void_p(Cd.data_ptr()), ipiv)
    clocks = ClockSampler(0)
    clocks.start()
    times, launches = [], 0
    for i in range(args.warmup + args.steps):
        if inplace_a:
            A.copy_(src)
        torch.cuda.synchronize()
        if i == args.warmup:
            clocks.mark()
        l0 = eigen_b200.kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call()
        e1.record()
        torch.cuda.synchronize()
        if i >= args.warmup:
            times.append(e0.elapsed_time(e1))
            launches = eigen_b200.kernel_launches() - l0
    clk = clocks.stop()
    ms = sum(times) / len(times)
    fl = level3_flops(r, n)
    variant = eigen_b200.last_variant()
    # e2e: the same F77 call on pinned HOST operands
    hA, hB, hC = src.cpu().pin_memory(), Bd.cpu().pin_memory(), Cd.cpu().pin_memory()
    del A, Bd, Cd, src
    torch.cuda.empty_cache()
    hA0 = hA.clone() if inplace_a else None
    hcall = level3_call(L, t, r, n, C.c_void_p(hA.data_ptr()), C.c_void_p(hB.data_ptr()), C.c_void_p(hC.data_ptr()), ipiv)
    hcall()
    e2e = []
    for _ in range(2):
        if inplace_a:
            hA.copy_(hA0)
        t0 = time.perf_counter()
        hcall()
        e2e.append((time.perf_counter() - t0) * 1e3)
    h2d, d2h = C.c_uint64(), C.c_uint64()
    L.b200blas_last_transfer(C.byref(h2d), C.byref(d2h))
    del hA, hB, hC, hA0
    try:
        cpu, _ = level3_cpu_reference(t, r, 2048)
    except Exception as e:  # reported, never required
        cpu = {"value": None, "unit": "TFLOP/s", "cores": 1, "kind": "reference", "sample": "failed: %r" % (e,)}
    issued = fl * (3.0 if t == "s" else 1.0)
    print(json.dumps({
        "metric": metric, "value": fl / (ms * 1e-3) / 1e12, "unit": "TFLOP/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": dtype + ("" if t == "d" else " (3xTF32 products, fp32 leaves)"), "data": "synthetic", "config": config,
        "gpu_launches": launches * args.steps, "kernel": variant, "clocks": clk,
        "e2e": {"value": fl / (min(e2e) * 1e-3) / 1e12, "unit": "TFLOP/s", "ms_per_step": min(e2e), "h2d_bytes_per_step": h2d.value,
                "d2h_bytes_per_step": d2h.value, "api": "%s%s_ (F77 C ABI) on pinned host operands" % (t, r)},
        "roofline": {"bound": "tensor", "achieved": issued / (ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                     "frac": issued / (ms * 1e-3) / 1e12 / peak if peak > 0 else None, "traffic": None, "kernel": variant,
                     "launches_per_step": launches, "algorithmic_flops_per_step": fl,
                     "note": "composite routine: the products run on the dgemm / sgemm kernels, the rest is the leaf chain (DESIGN.md 3b)"},
        "cpu_baseline": cpu}))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="dgemm16384", choices=sorted(WORKLOADS) + sorted(LEVEL3_WORKLOADS))
    ap.add_argument("--no-configs", action="store_true", help="skip the C1/C3/C4/C5 side configs of the default line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.workload in LEVEL3_WORKLOADS:
        return run_level3(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
