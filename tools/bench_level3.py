"""Measurement of the SURVEY 8 (f) rows, same shape as bench.py's line: one JSON line per routine with
  value     device-resident TFLOP/s (operands in HBM, F77 entry on device pointers, CUDA events on the legacy stream),
  e2e       the same call on pinned HOST operands (wall clock, copies inside),
  roofline  achieved / measured pipe peak (FP64 DMMA for d, TF32 tcgen05 for s with the 3x issue factor),
  cpu_baseline  the reference's own routine (oracle/_ref, single-threaded like blas/ and lapack/) on a bounded size.
usage: python tools/bench_level3.py [--n 8192] [--routines dsyrk,dtrsm,...] > profiles/bench_r01/level3.jsonl"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import eigen_b200  # noqa: E402
import oracle_api as oa  # noqa: E402

DT = {"s": torch.float32, "d": torch.float64}
NP = {"s": np.float32, "d": np.float64}


def flops_of(r, n):
    return {"syrk": n ** 3 * 1.0, "trsm": n ** 3 * 1.0, "trmm": n ** 3 * 1.0, "symm": 2.0 * n ** 3, "syr2k": 2.0 * n ** 3,
            "potrf": n ** 3 / 3.0, "getrf": 2.0 * n ** 3 / 3.0}[r]


def make_call(lib, t, r, n, pa, pb, pc, ipiv):
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--cpu-n", type=int, default=2048)
    ap.add_argument("--routines", default="dsyrk,dtrsm,dtrmm,dsymm,dsyr2k,dpotrf,dgetrf,ssyrk,strsm,ssymm,spotrf,sgetrf")
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    L = eigen_b200.require_device()
    eigen_b200.pipe_peak(0, 1500)   # ramp the clocks before the denominators are measured
    peaks = {"d": max(eigen_b200.pipe_peak(0, 800), eigen_b200.pipe_peak(0, 800)),
             "s": max(eigen_b200.pipe_peak(3, 800), eigen_b200.pipe_peak(3, 800))}
    for name in args.routines.split(","):
        t, r, n = name[0], name[1:], args.n
        g = torch.Generator(device="cuda").manual_seed(7)
        M = torch.rand(n, n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
        if r == "potrf":
            src = (M @ M.T / n + torch.eye(n, dtype=torch.float64, device="cuda")).to(DT[t])
        elif r in ("trsm", "trmm"):
            src = (M * (2.0 / n) + torch.eye(n, dtype=torch.float64, device="cuda") * 1.5).to(DT[t])
        else:
            src = M.to(DT[t])
        Bd = (torch.rand(n, n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1).to(DT[t])
        Cd = torch.ones(n, n, dtype=DT[t], device="cuda")
        del M
        ipiv = np.zeros(n, dtype=np.int32)
        inplace_a = r in ("potrf", "getrf")
        A = src.clone()
        call = make_call(L, t, r, n, C.c_void_p(A.data_ptr()), C.c_void_p(Bd.data_ptr()), C.c_void_p(Cd.data_ptr()), ipiv)
        times = []
        launches = 0
        for i in range(args.steps + 1):
            if inplace_a:
                A.copy_(src)
            torch.cuda.synchronize()
            l0 = eigen_b200.kernel_launches()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            call()
            e1.record()
            torch.cuda.synchronize()
            launches = eigen_b200.kernel_launches() - l0
            if i:
                times.append(e0.elapsed_time(e1))
        ms = sum(times) / len(times)
        fl = flops_of(r, n)
        value = fl / (ms * 1e-3) / 1e12
        variant = eigen_b200.last_variant()
        # e2e on pinned host operands
        hA, hB, hC = src.cpu().pin_memory(), Bd.cpu().pin_memory(), Cd.cpu().pin_memory()
        del A, Bd, Cd, src
        torch.cuda.empty_cache()
        hA0 = hA.clone() if inplace_a else None
        hcall = make_call(L, t, r, n, C.c_void_p(hA.data_ptr()), C.c_void_p(hB.data_ptr()), C.c_void_p(hC.data_ptr()), ipiv)
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
        # CPU baseline: the reference's routine, single thread, bounded size
        cn = args.cpu_n
        rng = np.random.default_rng(7)
        m = rng.uniform(-1, 1, (cn, cn))
        if r == "potrf":
            a = np.asfortranarray((m @ m.T / cn + np.eye(cn)).astype(NP[t]))
        elif r in ("trsm", "trmm"):
            a = np.asfortranarray((m * (2.0 / cn) + 1.5 * np.eye(cn)).astype(NP[t]))
        else:
            a = np.asfortranarray(m.astype(NP[t]))
        b = np.asfortranarray(rng.uniform(-1, 1, (cn, cn)).astype(NP[t]))
        c = np.ones((cn, cn), dtype=NP[t], order="F")
        cpiv = np.zeros(cn, dtype=np.int32)
        cpu = {"value": None, "unit": "TFLOP/s", "cores": 1, "kind": "reference", "sample": "unavailable"}
        try:
            lib = oa.ref_lapack() if inplace_a else oa.ref_blas()
            ccall = make_call(lib, t, r, cn, oa._ptr(a), oa._ptr(b), oa._ptr(c), cpiv)
            t0 = time.perf_counter()
            ccall()
            dt_s = time.perf_counter() - t0
            cpu = {"value": flops_of(r, cn) / dt_s / 1e12, "unit": "TFLOP/s", "cores": 1, "kind": "reference",
                   "sample": "%s%s_ of oracle/_ref (the reference's blas/ and lapack/ are single-threaded) at n=%d" % (t, r, cn)}
        except Exception as e:  # reported, never required
            cpu["sample"] = "failed: %r" % (e,)
        issued = fl * (3.0 if t == "s" else 1.0)
        line = {"metric": "%s%s TFLOP/s at n=%d" % (t, r, n), "value": value, "unit": "TFLOP/s", "n_gpus": 1, "steps": args.steps,
                "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "dtype": "f64" if t == "d" else "f32 (3xTF32)",
                "data": "synthetic", "config": {"workload": "%s%s_ order %d, side/uplo/trans = L/L/N, device-resident" % (t, r, n)},
                "e2e": {"value": fl / (min(e2e) * 1e-3) / 1e12, "unit": "TFLOP/s", "ms_per_step": min(e2e),
                        "h2d_bytes_per_step": h2d.value, "d2h_bytes_per_step": d2h.value, "api": "%s%s_ on pinned host operands" % (t, r)},
                "roofline": {"bound": "tensor", "achieved": issued / (ms * 1e-3) / 1e12, "peak": peaks[t], "unit": "TFLOP/s",
                             "frac": issued / (ms * 1e-3) / 1e12 / peaks[t], "traffic": None, "kernel": variant,
                             "launches_per_step": launches, "algorithmic_flops_per_step": fl},
                "cpu_baseline": cpu, "gpu_launches": launches * args.steps}
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
