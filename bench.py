#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for the Eigen dense-product hot path on B200.

Metric (BASELINE.json): dgemm TFLOP/s at N = 16384 (`C += A*B`, alpha = beta = 1, uniform[-1,1] inputs, C = 1 as in
bench/bench_gemm.cpp:130-133,211-214), 2*m*n*k flops per step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload dgemm16384|...]

* ours, N = 1: `value` = device-resident throughput (operands in HBM, one b200blas_gemm_dev call per step, CUDA
  events on the launching stream); `e2e` = the same product through the drop-in F77 entry point dgemm_ with HOST
  (pinned) operands, host<->device copies inside the timed region; `roofline` = FP64 DMMA pipe (tensor bound);
  `cpu_baseline` = the reference's OpenMP gebp path (oracle/_ref) on this box's host cores on a bounded slab.
* ours, N > 1 (torchrun, one rank per GPU): the product is partitioned by 2-D tiles of C (eigen_b200.parallelize),
  panels broadcast over NCCL, C tiles gathered to rank 0; strong scaling (total work fixed).
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

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []
        self.mark_at = 0

    def mark(self):
        """Call at the start of the timed region: only samples taken after this point are reported (the sampler is
        started before the warm-up because nvidia-smi needs ~0.3 s to deliver its first line)."""
        self.mark_at = len(self.lines)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
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

    def fill(r, c):
        x = np.empty((r, c), dtype=dt, order="F")
        blk = 1 << 22
        flat = x.reshape(-1, order="F")
        for i in range(0, flat.size, blk):
            j = min(flat.size, i + blk)
            if t in "cz":
                flat[i:j] = (rng.uniform(-1, 1, j - i) + 1j * rng.uniform(-1, 1, j - i)).astype(dt)
            else:
                flat[i:j] = rng.uniform(-1, 1, j - i).astype(dt)
        return x

    A = fill(m, k)
    B = fill(k, ns)
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
    sample = "%sgemm %dx%dx%d column slab (%d of %d columns) of the workload, C+=A*B, OMP threads=%d" % (t, m, ns, k, ns, n, cores)
    return flops / dt_s / 1e12, cores, sample, kind, dt_s * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    t, m, n, k, alpha, beta = WORKLOADS[args.workload]
    # bounded sample: ~3 s of CPU work per step
    target_flops = 0.5e12
    slab = max(256, min(n, int(target_flops / (FLOP_FACTOR[t] * m * k)) // 256 * 256))
    tf, cores, sample, kind, ms = cpu_reference_leg(t, m, n, k, alpha, beta, args.steps, args.warmup, slab)
    line = {
        "impl": "reference", "metric": metric_name(args.workload), "value": tf, "unit": "TFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": DTYPE_NAME[t].split()[0],
        "data": "synthetic", "config": workload_config(args.workload, args.gpus),
        "cpu_baseline": {"value": tf, "unit": "TFLOP/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": tf, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def metric_name(workload):
    t, m, n, k = WORKLOADS[workload][:4]
    if workload == "dgemm16384":
        return "dgemm TFLOP/s at N=16384"
    return "%sgemm TFLOP/s at %dx%dx%d" % (t, m, n, k)


def workload_config(workload, gpus):
    t, m, n, k, alpha, beta = WORKLOADS[workload]
    return {"workload": "%s: %sgemm C(%dx%d) = %g*A(%dx%d)*B(%dx%d) + %g*C, column-major, ld=rows, uniform[-1,1], C=1"
            % (workload, t, m, n, alpha, m, k, k, n, beta),
            "l2": ("inputs (%.2f GB) are larger than L2 (126 MB); no explicit flush" if ESIZE[t] * (m * k + k * n + m * n) > 2.5e8
                   else "inputs (%.2f GB) fit in L2: parity/diagnostic workload, not a bench line") % (ESIZE[t] * (m * k + k * n + m * n) / 1e9),
            "parallelism": "1 GPU" if gpus == 1 else "2-D tiles of C over %d GPUs (NCCL panel broadcast, C gather to rank 0)" % gpus}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import eigen_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)  # NCCL kernels must not queue behind GEMM CTAs
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), pg_options=opts)
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

    clocks = ClockSampler(local_rank)
    result = {}
    if world == 1:
        A, B = urand(m, k), urand(k, n)
        Cd = torch.ones(n, m, dtype=dt, device="cuda")
        stream = torch.cuda.current_stream()

        def step():
            r = eigen_b200.gemm_dev(t, "N", "N", m, n, k, alpha, A, m, B, k, beta, Cd, m, stream=stream.cuda_stream)
            assert r == 0, eigen_b200.last_error()

        clocks.start()
        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        clocks.mark()
        launches0 = eigen_b200.kernel_launches()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        ev[0].record(stream)
        for i in range(args.steps):
            step()
            ev[i + 1].record(stream)
        torch.cuda.synchronize()
        launches = eigen_b200.kernel_launches() - launches0
        total_ms = ev[0].elapsed_time(ev[-1])
        per = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
        result["clocks"] = clocks.stop()
        ms = total_ms / args.steps
        value = flops / (ms * 1e-3) / 1e12
        variant = eigen_b200.last_variant()
        # sanity: sampled rows against a float64 torch matmul of the same operands (not the oracle; a cheap guard)
        del ev
        # ---- roofline of the dominant kernel (one launch per step) ------------------------------------------
        pipe = 0 if t in "dz" else 3
        peak = eigen_b200.pipe_peak(pipe, 1500)
        peak_src = ("b200blas_pipe_peak(%s) measured in this run on this GPU; MEASURED_PEAKS.json has no %s figure"
                    % ("FP64 DMMA mma.sync.m8n8k4" if pipe == 0 else "TF32 tcgen05.mma", "FP64" if pipe == 0 else "TF32"))
        if peak <= 0 and pipe == 3:
            peak = 1100.0
            peak_src = "nominal dense TF32 1.1 PFLOP/s (fallback; microbenchmark unavailable)"
        issued = flops * (3.0 if t in "sc" and "tf32" in variant else 1.0)
        kern_ms = min(per)
        avg_ms = sum(per) / len(per)
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic_%s.json" % args.workload)
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_step")
            except Exception:
                traffic = None
        per_step = max(1, launches // max(args.steps, 1))
        result["roofline"] = {
            "bound": "tensor", "achieved": issued / (avg_ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
            "frac": issued / (avg_ms * 1e-3) / 1e12 / peak if peak > 0 else None, "traffic": traffic,
            "kernel": variant, "step_ms_avg": avg_ms, "step_ms_best": kern_ms, "launches_per_step": per_step,
            "note": "one step = one product = %d launch(es) of the kernel (k-slices / pack + product); achieved, traffic "
                    "and the algorithmic figures are per step" % per_step,
            "algorithmic_flops_per_step": flops, "tensor_pipe_flops_per_step": issued,
            "algorithmic_bytes_per_step": ESIZE[t] * (m * k + k * n + 2 * m * n), "peak_source": peak_src,
        }
        # ---- e2e: the drop-in F77 entry point with HOST operands --------------------------------------------
        del A, B, Cd
        torch.cuda.empty_cache()
        npdt = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}[t]
        hA = torch.empty(k, m, dtype=dt).pin_memory()
        hB = torch.empty(n, k, dtype=dt).pin_memory()
        hC = torch.empty(n, m, dtype=dt).pin_memory()
        rng = np.random.default_rng(42)
        for h in (hA, hB):
            v = h.numpy().reshape(-1)
            blk = 1 << 24
            for i in range(0, v.size, blk):
                j = min(v.size, i + blk)
                if t in "cz":
                    v[i:j] = (rng.uniform(-1, 1, j - i) + 1j * rng.uniform(-1, 1, j - i)).astype(npdt)
                else:
                    v[i:j] = rng.uniform(-1, 1, j - i).astype(npdt)
        hC.fill_(1)
        fn = getattr(L, t + "gemm_")
        al = np.array([alpha], dtype=npdt)
        be = np.array([beta], dtype=npdt)
        ints = [C.c_int(x) for x in (m, n, k, m, k, m)]

        def e2e_step():
            r = fn(b"N", b"N", C.byref(ints[0]), C.byref(ints[1]), C.byref(ints[2]), al.ctypes.data_as(C.c_void_p),
                   C.c_void_p(hA.data_ptr()), C.byref(ints[3]), C.c_void_p(hB.data_ptr()), C.byref(ints[4]),
                   be.ctypes.data_as(C.c_void_p), C.c_void_p(hC.data_ptr()), C.byref(ints[5]))
            assert r == 0, eigen_b200.last_error()

        e2e_steps = max(1, min(args.steps, 3))
        e2e_step()  # warm-up (allocates the staging buffers)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) / e2e_steps * 1e3
        h2d, d2h = C.c_uint64(), C.c_uint64()
        L.b200blas_last_transfer(C.byref(h2d), C.byref(d2h))
        result["e2e"] = {"value": flops / (e2e_ms * 1e-3) / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": h2d.value,
                         "d2h_bytes_per_step": d2h.value, "ms_per_step": e2e_ms, "steps": e2e_steps,
                         "api": "%sgemm_ (F77 C ABI, include/b200blas.h) on pinned host operands" % t}
        del hA, hB, hC
        # ---- CPU baseline: the reference's OpenMP gebp on this box's cores, bounded slab ---------------------
        try:
            slab = max(256, min(n, int(1.0e12 / (FLOP_FACTOR[t] * m * k)) // 256 * 256))
            tf, cores, sample, kind, _ = cpu_reference_leg(t, m, n, k, alpha, beta, 2, 1, slab)
            result["cpu_baseline"] = {"value": tf, "unit": "TFLOP/s", "cores": cores, "kind": kind, "sample": sample}
        except Exception as e:  # the baseline is reported, never required for the GPU number
            result["cpu_baseline"] = {"value": None, "unit": "TFLOP/s", "cores": os.cpu_count(), "kind": "port",
                                      "sample": "failed: %r" % (e,)}
        ms_per_step, n_launch = ms, launches
    else:
        from eigen_b200 import parallelize
        job = parallelize.DistGemm(t, m, n, k, alpha, beta)
        if rank == 0:
            A, B = urand(m, k), urand(k, n)
            Cd = torch.ones(n, m, dtype=dt, device="cuda")
        else:
            A = B = Cd = None
        clocks.start()
        for _ in range(args.warmup):
            job.run(A, B, Cd)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        clocks.mark()
        launches0 = eigen_b200.kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            job.run(A, B, Cd)
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        launches = eigen_b200.kernel_launches() - launches0
        tms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ln = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(ln, op=dist.ReduceOp.SUM)
        result["clocks"] = clocks.stop()
        ms_per_step = tms.item() / args.steps
        value = flops / (ms_per_step * 1e-3) / 1e12
        n_launch = int(ln.item())
        variant = eigen_b200.last_variant()
        del A, B, Cd
        torch.cuda.empty_cache()
        # ---- e2e at N GPUs: operands in the caller's HOST memory (POSIX shm, mapped + pinned by every rank);
        # every GPU pulls its share over its own PCIe link, A is all-gathered over NVLink, C_j goes straight back.
        try:
            tag = "b200bench_%s_" % os.environ.get("MASTER_PORT", "0")
            hA = parallelize.shared_host_tensor(tag + "A", (k, m), dt, rank == 0) if rank == 0 else None
            hB = parallelize.shared_host_tensor(tag + "B", (n, k), dt, rank == 0) if rank == 0 else None
            hC = parallelize.shared_host_tensor(tag + "C", (n, m), dt, rank == 0) if rank == 0 else None
            if rank == 0:
                g = torch.Generator().manual_seed(42)
                for h in (hA, hB):
                    flat = h.view(-1)
                    for i0 in range(0, flat.numel(), 1 << 24):
                        seg = flat[i0:i0 + (1 << 24)]
                        if t in "cz":
                            torch.view_as_real(seg).uniform_(-1, 1, generator=g)
                        else:
                            seg.uniform_(-1, 1, generator=g)
                hC.fill_(1)
            dist.barrier()
            if rank != 0:
                hA = parallelize.shared_host_tensor(tag + "A", (k, m), dt, False)
                hB = parallelize.shared_host_tensor(tag + "B", (n, k), dt, False)
                hC = parallelize.shared_host_tensor(tag + "C", (n, m), dt, False)
            if k % world == 0:
                kb = k // world
                parallelize.pin_host_range(hA[rank * kb:(rank + 1) * kb])
            else:
                parallelize.pin_host_range(hA)
            if job.nj > 0:
                parallelize.pin_host_range(hB[job.c0:job.c1])
                parallelize.pin_host_range(hC[job.c0:job.c1])
            e2e_steps = max(1, min(args.steps, 3))
            job.run_host(hA, hB, hC)  # warm-up
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                job.run_host(hA, hB, hC)
            torch.cuda.synchronize()
            dist.barrier()
            el = torch.tensor([(time.perf_counter() - t0) / e2e_steps * 1e3], dtype=torch.float64, device="cuda")
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
            bts = torch.tensor([getattr(job, "h2d_bytes", 0), getattr(job, "d2h_bytes", 0)], dtype=torch.int64, device="cuda")
            dist.all_reduce(bts, op=dist.ReduceOp.SUM)
            result["e2e"] = {"value": flops / (el.item() * 1e-3) / 1e12, "unit": "TFLOP/s",
                             "h2d_bytes_per_step": int(bts[0].item()), "d2h_bytes_per_step": int(bts[1].item()),
                             "ms_per_step": el.item(), "steps": e2e_steps,
                             "api": "DistGemm.run_host: A/B/C in shared pinned host memory, each GPU loads its share "
                                    "over its own PCIe link, A all-gathered over NVLink, C tiles written back per rank"}
            dist.barrier()
            if rank == 0:
                for nm in ("A", "B", "C"):
                    try:
                        os.unlink("/dev/shm/" + tag + nm)
                    except OSError:
                        pass
        except Exception as e:  # never lose the device-resident number to a host-memory problem
            result["e2e"] = {"value": None, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                             "error": repr(e)[:300]}
        result["residency"] = "A, B, C resident on rank 0 (root-resident); panels broadcast, C tiles gathered inside the timed region"
    if rank == 0:
        line = {
            "metric": metric_name(args.workload), "value": value, "unit": "TFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": DTYPE_NAME[t].split()[0], "data": "synthetic",
            "config": workload_config(args.workload, world), "gpu_launches": n_launch, "kernel": variant,
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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.workload in LEVEL3_WORKLOADS:
        return run_level3(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
