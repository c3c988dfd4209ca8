"""Device-resident timing of ?syrk_/?herk_ (F77 entry on device pointers) next to the matching gemm.  Usage: time_rankk.py [n] [k]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import eigen_b200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
k = int(sys.argv[2]) if len(sys.argv) > 2 else n
L = eigen_b200.require_device()
DT = {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128}
for name in ("dsyrk_", "zherk_", "ssyrk_", "cherk_"):
    t = name[0]
    nn, kk = (n, k) if t in "sd" else (n // 2, k // 2)
    A = torch.rand(kk, nn, dtype=torch.float64, device="cuda").to(DT[t])
    Cm = torch.zeros(nn, nn, dtype=DT[t], device="cuda")
    herk = "herk" in name
    rt = C.c_float if t in "sc" else C.c_double
    if herk or t in "sd":
        al, be = rt(1.0), rt(0.0)
    else:
        al, be = (rt * 2)(1.0, 0.0), (rt * 2)(0.0, 0.0)
    ints = [C.c_int(v) for v in (nn, kk, nn, nn)]
    fn = getattr(L, name)

    def call():
        return fn(b"L", b"N", C.byref(ints[0]), C.byref(ints[1]), C.byref(al), C.c_void_p(A.data_ptr()), C.byref(ints[2]),
                  C.byref(be), C.c_void_p(Cm.data_ptr()), C.byref(ints[3]))
    for _ in range(2):
        assert call() == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    mul = 4 if t in "cz" else 1
    useful = mul * nn * (nn + 1) * kk / (ms * 1e-3) / 1e12      # flops of the referenced triangle
    print("%s n=%d k=%d  %.2f ms  %.2f TFLOP/s (triangle flops)  gemm-equivalent %.2f  variant %s" %
          (name, nn, kk, ms, useful, 2 * useful, eigen_b200.last_variant()), flush=True)
