"""Python mirror of the multi-GPU partitioner -- the B200 counterpart of ``parallelize_gemm``
(``Eigen/src/Core/products/Parallelizer.h:85-157``).

The partitioner itself lives BEHIND the C ABI (``eigen_b200/csrc/multi.cu``, ``include/b200blas.h`` section 3): one
process drives N GPUs, so ``dgemm_`` from an unmodified ``EIGEN_USE_BLAS`` program uses all of them
(``B200BLAS_NGPUS=N``).  Round 1's torch.distributed/NCCL driver (one process per GPU, reachable from Python only) is
gone; this module only exposes the plan for inspection and a convenience wrapper.  No arithmetic happens here.
"""
import eigen_b200


def grid_for(world):
    """pr x pc grid of C tiles the library uses for `world` devices (2 -> 1x2, 4 -> 2x2, 8 -> 2x4)."""
    info, _ = eigen_b200.multi_plan("d", "N", "N", 4096, 4096, 4096, 1.0, 1.0, world, cap=0)
    return info.pr, info.pc


def partition(m, n, world, grid=(0, 0)):
    """Tile of C owned by each device: list of (r0, r1, c0, c1), device = i * pc + j."""
    info, _ = eigen_b200.multi_plan("d", "N", "N", m, n, 1024, 1.0, 1.0, world, grid=grid, cap=0)
    rows, cols = list(info.row_cut), list(info.col_cut)
    return [(rows[i], rows[i + 1], cols[j], cols[j + 1]) for i in range(info.pr) for j in range(info.pc)]


def traffic(t, m, n, k, world, beta=1.0, host_origin=False, grid=(0, 0)):
    """Bytes the plan moves: over the slow link (origin -> device), between peers, and back to the origin."""
    es = {"s": 4, "d": 8, "c": 8, "z": 16}[t]
    info, steps = eigen_b200.multi_plan(t, "N", "N", m, n, k, 1.0, beta, world, grid=grid, host_origin=host_origin)
    out = {"from_origin": 0, "peer_to_peer": 0, "to_origin": 0, "grid": (info.pr, info.pc), "k_chunks": info.nchunks, "steps": len(steps)}
    for s in steps:
        if s.kind != 0:
            continue
        nbytes = s.x.rows * s.x.cols * es
        if s.x.loc < 0:
            out["from_origin"] += nbytes
        elif s.z.loc < 0:
            out["to_origin"] += nbytes
        else:
            out["peer_to_peer"] += nbytes
    return out


def gemm_multi(t, transa, transb, m, n, k, alpha, dA, lda, dB, ldb, beta, dC, ldc, ndev, stream=None):
    """One product on `ndev` GPUs, operands resident on the current device; asynchronous on `stream`."""
    eigen_b200.set_devices(ndev)
    return eigen_b200.gemm_dev(t, transa, transb, m, n, k, alpha, dA, lda, dB, ldb, beta, dC, ldc, stream=stream)
