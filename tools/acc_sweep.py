"""Accuracy of the float kernels against an fp64 reference as k grows (device-side check, torch float64 matmul).
Prints relative Frobenius error in units of eps and the worst netlib gauge ratio |c-ref|/(eps*sum|a||b|)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import eigen_b200  # noqa: E402

eps = 2.0 ** -23
m = n = 1024
for variant in sys.argv[1:] or ["tf32x3", "simt"]:
    for k in (256, 1024, 4096, 8192, 16384):
        g = torch.Generator(device="cuda").manual_seed(k)
        A = torch.rand(k, m, dtype=torch.float32, device="cuda", generator=g) * 2 - 1
        B = torch.rand(n, k, dtype=torch.float32, device="cuda", generator=g) * 2 - 1
        C = torch.zeros(n, m, dtype=torch.float32, device="cuda")
        assert eigen_b200.gemm_dev("s", "N", "N", m, n, k, 1.0, A, m, B, k, 0.0, C, m, variant=variant) == 0
        ref = B.double() @ A.double()
        gauge = B.double().abs() @ A.double().abs()
        err = (C.double() - ref)
        fro = (err.norm() / ref.norm()).item()
        ratio = (err.abs() / (eps * gauge)).max().item()
        bias = (err * ref.sign()).mean().item() / (eps * gauge.mean().item())
        print("%-7s kchunk=%s k=%5d  rel_fro/eps %8.2f  (k*eps bound %6d)  max gauge ratio %7.3f  signed bias/(eps*G) %+.3f  [%s]" % (
            variant, os.environ.get("B200BLAS_TF32_KCHUNK", "default"), k, fro / eps, k, ratio, bias, eigen_b200.last_variant()))
