"""A/B timing of one device-resident dgemm through an explicitly given build of the library.
usage: python tools/ab_lib.py <path/to/libb200blas.so> [n] [iters]"""
import ctypes as C
import sys

import torch

import os
lib = C.CDLL(os.path.abspath(sys.argv[1]))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
i, vp = C.c_int, C.c_void_p
lib.b200blas_gemm_dev.argtypes = [i, C.c_char, C.c_char, i, i, i, vp, vp, C.c_int64, vp, C.c_int64, vp, vp, C.c_int64, vp, i]
A = torch.rand(n, n, dtype=torch.float64, device="cuda") * 2 - 1
B = torch.rand(n, n, dtype=torch.float64, device="cuda") * 2 - 1
Cm = torch.ones(n, n, dtype=torch.float64, device="cuda")
one = C.c_double(1.0)
s = torch.cuda.current_stream().cuda_stream


def run():
    return lib.b200blas_gemm_dev(1, b"N", b"N", n, n, n, C.byref(one), A.data_ptr(), n, B.data_ptr(), n, C.byref(one), Cm.data_ptr(), n, s, 0)


for _ in range(2):
    assert run() == 0
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
ev[0].record()
for k in range(iters):
    run()
    ev[k + 1].record()
torch.cuda.synchronize()
per = [ev[k].elapsed_time(ev[k + 1]) for k in range(iters)]
print("%s n=%d ms/step %s best %.2f TF" % (sys.argv[1].split("/")[-1], n, ["%.2f" % x for x in per], 2.0 * n ** 3 / (min(per) * 1e-3) / 1e12))
