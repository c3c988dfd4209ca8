"""Host logic of the multi-GPU partitioner on CPU: world_size-2 (and 4) gloo process groups, with the device
kernel replaced by a float64 matmul stand-in inside the TEST processes only (the product has no CPU path)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eigen_b200 import parallelize  # noqa: E402


def test_split_and_partition_follow_parallelize_gemm_rules():
    # Parallelizer.h:140-151 -- equal blocks rounded to the register/tile quantum, the last takes the remainder
    assert parallelize.split(16384, 8, 128) == [(i * 2048, (i + 1) * 2048) for i in range(8)]
    s = parallelize.split(1000, 4, 128)
    assert s == [(0, 256), (256, 512), (512, 768), (768, 1000)]
    assert parallelize.split(100, 4, 128) == [(0, 100), (100, 100), (100, 100), (100, 100)]
    tiles = parallelize.partition(16384, 16384, 8, grid=(2, 4))
    assert tiles[0] == (0, 8192, 0, 4096) and tiles[7] == (8192, 16384, 12288, 16384)
    # the tiles cover C exactly once
    cover = np.zeros((300, 500), dtype=int)
    for (r0, r1, c0, c1) in parallelize.partition(300, 500, 4, grid=(2, 2), quantum=128):
        cover[r0:r1, c0:c1] += 1
    assert np.all(cover == 1)
    assert parallelize.chunk_ranges(1000, 3) == [(0, 512), (512, 1000)]
    assert parallelize.chunk_ranges(16384, 0) == [(0, 1024), (1024, 2048), (2048, 4096), (4096, 8192), (8192, 16384)]
    assert parallelize.chunk_ranges(1000, 0) == [(0, 1000)]
    assert parallelize.subslab_ranges(2048, 0) == [(0, 1024), (1024, 1792), (1792, 2048)]
    assert parallelize.subslab_ranges(300, 0) == [(0, 300)]
    assert parallelize.grid_for(8) == (1, 8)


def _cpu_gemm_stand_in(t, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, stream=None, variant="auto"):
    """TEST DOUBLE for eigen_b200.gemm_dev on CPU tensors (views of column-major storage, ld = stride(0))."""
    assert ta == "N" and tb == "N"
    a = torch.as_strided(A, (k, m), (lda, 1))      # (k, m) rows = columns of the column-major m x k
    b = torch.as_strided(B, (n, k), (ldb, 1))
    c = torch.as_strided(C, (n, m), (ldc, 1))
    prod = (b @ a) * alpha                           # (n, m) == (A*B)^T
    if beta == 0:
        c.copy_(prod)
    else:
        c.mul_(beta).add_(prod)
    return 0


def _worker(rank, world, port, grid, shape, beta, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import eigen_b200
    eigen_b200.gemm_dev = _cpu_gemm_stand_in   # patched in this test process only
    m, n, k = shape
    if k > 4000:
        job = parallelize.DistGemm("d", m, n, k, 0.7, beta, grid=grid)
    else:
        job = parallelize.DistGemm("d", m, n, k, 0.7, beta, kchunks=3, subslabs=2, grid=grid)
    if rank == 0:
        g = torch.Generator().manual_seed(5)
        A = torch.rand(k, m, dtype=torch.float64, generator=g) * 2 - 1
        B = torch.rand(n, k, dtype=torch.float64, generator=g) * 2 - 1
        C = torch.rand(n, m, dtype=torch.float64, generator=g)
        want = beta * C + 0.7 * (B @ A)
        for _ in range(2):  # the second run checks buffer reuse
            Cw = C.clone()
            job.run(A, B, Cw)
        err = (Cw - want).abs().max().item()
        out.put(err)
    else:
        for _ in range(2):
            job.run(None, None, None)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,grid,shape,beta", [
    (2, (1, 2), (300, 520, 700), 1.0),
    (2, (2, 1), (300, 260, 515), 0.0),
    (4, (2, 2), (390, 410, 600), 1.3),
    (4, (1, 4), (130, 1000, 300), 1.0),
    (2, (1, 2), (140, 4300, 4200), 1.0),   # default doubling k-chunks and 1/2 + 3/8 + 1/8 sub-slabs
])
def test_distgemm_gloo(world, grid, shape, beta):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, grid, shape, beta, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert out.get(timeout=5) < 1e-10
