#!/usr/bin/env python
"""Only the end-to-end legs of bench.py (dgemm_ on pageable and on page-locked host operands), for A/B runs of the host
pipeline:  [B200BLAS_COPY_THREADS=n] python tools/e2e_probe.py [--workload dgemm16384] [--steps 2] [--gpus N]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import eigen_b200  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="dgemm16384")
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--gpus", type=int, default=1)
a = ap.parse_args()
L = eigen_b200.require_device()
if a.gpus > 1:
    assert eigen_b200.set_devices(a.gpus) == a.gpus
t, m, n, k, alpha, beta = bench.WORKLOADS[a.workload]
e2e, e2e_pinned, hA, hB, hC, done = bench.e2e_legs(L, torch, eigen_b200, t, m, n, k, alpha, beta, a.steps, a.gpus)
chk = bench.sampled_row_check(t, m, n, k, alpha * done, 1.0, hA, hB, 1.0, hC[[0, m // 2, m - 1]], [0, m // 2, m - 1])
print(json.dumps({"workload": a.workload, "gpus": a.gpus, "copy_threads": os.environ.get("B200BLAS_COPY_THREADS", "default"),
                  "cores": os.cpu_count(), "e2e_pageable_tflops": e2e["value"], "e2e_pageable_ms": e2e["ms_per_step"],
                  "e2e_pinned_tflops": e2e_pinned["value"], "e2e_pinned_ms": e2e_pinned["ms_per_step"], "checked_ok": chk["ok"]}))
