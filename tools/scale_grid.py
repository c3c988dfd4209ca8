#!/usr/bin/env python
"""A/B table of partition grids for one device-resident product driven from ONE process through the C ABI
(b200blas_set_devices / b200blas_set_grid, include/b200blas.h section 3).  Prints one JSON line per grid:
    python tools/scale_grid.py --ndev 8 --grids 2x4,1x8,4x2 [--workload dgemm16384] [--steps 5]
CUDA events on the launching stream of the root device; the call returns to that stream only after every device joined."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import eigen_b200  # noqa: E402
from bench import WORKLOADS, FLOP_FACTOR, torch_dtype  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ndev", type=int, default=8)
ap.add_argument("--grids", default="2x4,1x8")
ap.add_argument("--workload", default="dgemm16384")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=3)
a = ap.parse_args()
L = eigen_b200.require_device()
t, m, n, k, alpha, beta = WORKLOADS[a.workload]
dt = torch_dtype(t)
torch.cuda.set_device(0)
g = torch.Generator(device="cuda").manual_seed(1)
A = torch.rand(k, m, dtype=dt, device="cuda", generator=g) * 2 - 1
B = torch.rand(n, k, dtype=dt, device="cuda", generator=g) * 2 - 1
Cd = torch.ones(n, m, dtype=dt, device="cuda")
st = torch.cuda.current_stream()
flops = FLOP_FACTOR[t] * m * n * k
assert eigen_b200.set_devices(a.ndev) == a.ndev
for grid in a.grids.split(","):
    pr, pc = (int(x) for x in grid.split("x"))
    assert L.b200blas_set_grid(pr, pc) == 0

    def call():
        r = eigen_b200.gemm_dev(t, "N", "N", m, n, k, alpha, A, m, B, k, beta, Cd, m, stream=st.cuda_stream)
        assert r == 0, eigen_b200.last_error()
    for _ in range(a.warmup):
        call()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    ev[0].record(st)
    for i in range(a.steps):
        call()
        ev[i + 1].record(st)
    torch.cuda.synchronize()
    per = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.steps)]
    ms = sum(per) / len(per)
    print(json.dumps({"workload": a.workload, "ndev": a.ndev, "grid": grid, "ms_per_step": ms, "ms_best": min(per),
                      "tflops": flops / ms / 1e9, "steps": a.steps, "residency": "root-resident (A, B, C on GPU 0)"}), flush=True)
L.b200blas_set_grid(0, 0)
