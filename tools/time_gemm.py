"""Device-resident timing of one product (CUDA events), for variant / tile-configuration sweeps.
usage: python tools/time_gemm.py <type> <m> <n> <k> [ta tb [iters [variant]]]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import eigen_b200  # noqa: E402

t, m, n, k = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
ta = sys.argv[5] if len(sys.argv) > 5 else "N"
tb = sys.argv[6] if len(sys.argv) > 6 else "N"
iters = int(sys.argv[7]) if len(sys.argv) > 7 else 5
variant = sys.argv[8] if len(sys.argv) > 8 else "auto"
dt = {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128}[t]
ra, ca = (m, k) if ta == "N" else (k, m)
rb, cb = (k, n) if tb == "N" else (n, k)
A = torch.rand(ca, ra, dtype=dt, device="cuda") * 2 - 1
B = torch.rand(cb, rb, dtype=dt, device="cuda") * 2 - 1
C = torch.ones(n, m, dtype=dt, device="cuda")
run = lambda: eigen_b200.gemm_dev(t, ta, tb, m, n, k, 1.0, A, ra, B, rb, 1.0, C, m, variant=variant)  # noqa: E731
for _ in range(2):
    assert run() == 0, eigen_b200.last_error()
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
e[0].record()
for i in range(iters):
    run()
    e[i + 1].record()
torch.cuda.synchronize()
ms = [e[i].elapsed_time(e[i + 1]) for i in range(iters)]
fl = (8.0 if t in "cz" else 2.0) * m * n * k
print("%sgemm %s%s %dx%dx%d cfg=%s variant=%s best %.3f ms %.2f TF | avg %.3f ms %.2f TF" % (
    t, ta, tb, m, n, k, os.environ.get("B200BLAS_DMMA_CFG", "-"), eigen_b200.last_variant(), min(ms), fl / min(ms) / 1e9,
    sum(ms) / len(ms), fl / (sum(ms) / len(ms)) / 1e9))
