"""Device-resident timing of ?trsm_/?trmm_/?symm_/?syr2k_ through the F77 entries on device pointers.
Usage: time_level3.py [n]   (square operands of order n; complex types use n/2)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import eigen_b200  # noqa: E402

n0 = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
L = eigen_b200.require_device()
DT = {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128}


def scal(t, v):
    rt = C.c_float if t in "sc" else C.c_double
    return rt(v) if t in "sd" else (rt * 2)(v, 0.0)


def timed(call, reps=3):
    for _ in range(2):
        call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for t in "dzsc":
    n = n0 if t in "sd" else n0 // 2
    mul = 4 if t in "cz" else 1
    A = (torch.rand(n, n, dtype=torch.float64, device="cuda") * (2.0 / n)).to(DT[t])
    A += torch.eye(n, dtype=DT[t], device="cuda") * 1.5
    B = torch.rand(n, n, dtype=torch.float64, device="cuda").to(DT[t])
    Cm = torch.zeros(n, n, dtype=DT[t], device="cuda")
    ints = [C.c_int(n)] * 6
    i = [C.byref(C.c_int(n)) for _ in range(6)]
    one, zero = scal(t, 1.0), scal(t, 0.0)
    pa, pb, pc = C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), C.c_void_p(Cm.data_ptr())
    for nm, flops, call in [
        (t + "trmm_", mul * n ** 3, lambda: getattr(L, t + "trmm_")(b"L", b"L", b"N", b"N", i[0], i[1], C.byref(one), pa, i[2], pb, i[3])),
        (t + "trsm_", mul * n ** 3, lambda: getattr(L, t + "trsm_")(b"L", b"L", b"N", b"N", i[0], i[1], C.byref(one), pa, i[2], pb, i[3])),
        (t + "trsm_ R/L/T", mul * n ** 3, lambda: getattr(L, t + "trsm_")(b"R", b"L", b"T", b"N", i[0], i[1], C.byref(one), pa, i[2], pb, i[3])),
        (t + "symm_", mul * 2 * n ** 3, lambda: getattr(L, t + "symm_")(b"L", b"L", i[0], i[1], C.byref(one), pa, i[2], pb, i[3], C.byref(zero), pc, i[4])),
        (t + "syr2k_", mul * 2 * n ** 3, lambda: getattr(L, t + "syr2k_")(b"L", b"N", i[0], i[1], C.byref(one), pa, i[2], pb, i[3], C.byref(zero), pc, i[4])),
    ]:
        before = eigen_b200.kernel_launches()
        call()
        launches = eigen_b200.kernel_launches() - before
        ms = timed(call)
        print("%-14s n=%d  %8.2f ms  %7.2f TFLOP/s  (%d launches)" % (nm, n, ms, flops / (ms * 1e-3) / 1e12, launches), flush=True)
    # the LU panel shape: 256 x 256 unit-lower triangle against a long right-hand side (PartialPivLU.h:490)
    if t == "d":
        m, nr = 256, 16384
        Bp = torch.rand(nr, m, dtype=DT[t], device="cuda")
        im, inr = C.byref(C.c_int(m)), C.byref(C.c_int(nr))
        ms = timed(lambda: L.dtrsm_(b"L", b"L", b"N", b"U", im, inr, C.byref(one), pa, i[2], C.c_void_p(Bp.data_ptr()), im))
        print("dtrsm_ LU panel m=%d n=%d  %.3f ms  %.2f TFLOP/s" % (m, nr, ms, m * m * nr / (ms * 1e-3) / 1e12), flush=True)
