"""Host-side rank logic of bench.py at N > 1 on CPU: world_size-2 gloo process group.

The data path of the multi-GPU product is a single-process multi-device engine behind the C ABI (checked by
tests/test_multi_plan.py on the CPU and tests/test_gpu_multi.py on hardware); what torch.distributed carries in
bench.py is the bracket around the timed region -- barriers and the max-over-ranks reduction -- and the rule that only
rank 0 prints.  Those are exercised here without a GPU."""
import io
import json
import os
import socket
import sys
from contextlib import redirect_stdout

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    # rank 0 drives every device and reports the step time; the other ranks report 0 for that leg
    got = bench.max_over_ranks(12.5 if rank == 0 else 0.0, device="cpu")
    # the pre-sharded leg: every rank times its own tile, the slowest one counts
    got2 = bench.max_over_ranks(3.0 + rank, device="cpu")
    dist.barrier()
    out.put((rank, got, got2))
    dist.destroy_process_group()


def test_max_over_ranks_world_size_2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, 12.5, 4.0), (1, 12.5, 4.0)]


def test_reference_arm_only_rank_0_speaks(monkeypatch):
    import bench
    monkeypatch.setenv("RANK", "1")
    buf = io.StringIO()
    with redirect_stdout(buf):
        assert bench.run_reference(type("A", (), {"workload": "dgemm16384", "gpus": 2, "steps": 1, "warmup": 1})()) == 0
    assert buf.getvalue() == ""


def test_cpu_legs_share_one_slab_rule():
    """`cpu_baseline` and `--impl reference` must time the same sample (>= 256 columns per core)."""
    import bench
    for cores in (8, 16, 32, 128):
        cols = bench.cpu_slab_columns("d", 16384, 16384, 16384, cores)
        assert cols == min(16384, 256 * cores)
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count("= cpu_slab_columns(t, m, n, k") == 2 and "target_flops" not in src


def test_workload_config_names_the_grid_actually_used():
    import bench
    cfg = bench.workload_config("dgemm16384", 8)
    assert "2x4 grid" in cfg["parallelism"] and "B200BLAS_NGPUS=8" in cfg["parallelism"]
    assert bench.workload_config("dgemm16384", 4)["parallelism"].startswith("2x2 grid")
    assert bench.workload_config("dgemm16384", 1)["parallelism"] == "1 GPU"
    json.dumps(cfg)
