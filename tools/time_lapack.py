"""Device-resident timing of ?potrf_/?getrf_ through the F77 entries on device pointers.  Usage: time_lapack.py [n ...]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import eigen_b200  # noqa: E402

L = eigen_b200.require_device()
DT = {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128}
sizes = [int(x) for x in sys.argv[1:]] or [8192]
for n in sizes:
    for t in "ds":
        g = torch.Generator(device="cuda").manual_seed(1)
        M = (torch.rand(n, n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1)
        spd = (M @ M.T / n + torch.eye(n, dtype=torch.float64, device="cuda")).to(DT[t])
        gen = (M.clone()).to(DT[t])
        del M
        nn, info = C.c_int(n), C.c_int(0)
        ipiv = np.zeros(n, dtype=np.int32)
        for name, src, flops in (("potrf_", spd, n ** 3 / 3.0), ("getrf_", gen, 2.0 * n ** 3 / 3.0)):
            best = 1e30
            launches = 0
            for rep in range(3):
                A = src.clone()
                torch.cuda.synchronize()
                before = eigen_b200.kernel_launches()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if name == "potrf_":
                    getattr(L, t + name)(b"L", C.byref(nn), C.c_void_p(A.data_ptr()), C.byref(nn), C.byref(info))
                else:
                    getattr(L, t + name)(C.byref(nn), C.byref(nn), C.c_void_p(A.data_ptr()), C.byref(nn), ipiv.ctypes.data_as(C.POINTER(C.c_int)), C.byref(info))
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
                launches = eigen_b200.kernel_launches() - before
            print("%s%s n=%d  %9.2f ms  %7.2f TFLOP/s  info=%d  (%d launches)" % (t, name, n, best, flops / (best * 1e-3) / 1e12, info.value, launches), flush=True)
