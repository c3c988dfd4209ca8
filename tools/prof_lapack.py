"""One device-resident dpotrf or dgetrf of order n (for ncu launch lists).  usage: prof_lapack.py potrf|getrf n"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import eigen_b200  # noqa: E402

which, n = sys.argv[1], int(sys.argv[2])
L = eigen_b200.require_device()
M = torch.rand(n, n, dtype=torch.float64, device="cuda") * 2 - 1
A = (M @ M.T / n + torch.eye(n, dtype=torch.float64, device="cuda")) if which == "potrf" else M
nn, info = C.c_int(n), C.c_int(0)
ipiv = np.zeros(n, dtype=np.int32)
if which == "potrf":
    L.dpotrf_(b"L", C.byref(nn), C.c_void_p(A.data_ptr()), C.byref(nn), C.byref(info))
else:
    L.dgetrf_(C.byref(nn), C.byref(nn), C.c_void_p(A.data_ptr()), C.byref(nn), ipiv.ctypes.data_as(C.POINTER(C.c_int)), C.byref(info))
print(which, n, "info", info.value, "launches", eigen_b200.kernel_launches())
