"""Multi-GPU partitioner for the dense product -- the B200 counterpart of ``parallelize_gemm``.

The reference splits ``C`` into one column slab per OpenMP thread (multiples of ``nr = 4`` columns), lets every
thread read all of ``A`` and its own columns of ``B`` and packs ``A`` cooperatively
(``Eigen/src/Core/products/Parallelizer.h:85-157``, ``GeneralMatrixMatrix.h:83-152``).  Here the workers are GPUs
(one process each, ``torch.distributed`` over NCCL/NVLink): ``C`` is cut into a ``pr x pc`` grid of tiles, rank
``(i, j)`` receives the row panel ``A_i`` and the column panel ``B_j`` and owns ``C_ij``; ``k`` is never split, so
there is no reduction -- only panel distribution (broadcast / send) and the gather of the ``C`` tiles.

Residency model of :class:`DistGemm.run`: ``A``, ``B``, ``C`` live on rank 0 ("root-resident", like the caller's
matrices in the reference).  Pipeline on every rank:

* phase 1 -- ``k`` is streamed in chunks: chunk ``c+1`` of ``A_i``/``B_j`` travels over NVLink (comm stream) while the
  first column sub-slab of ``C_ij`` accumulates chunk ``c`` on the compute stream (``beta = 1`` after the first);
  panels stay resident in HBM;
* phase 2 -- the remaining column sub-slabs are single full-``k`` products; each finished sub-slab is sent to the root
  while the next one computes; the root folds ``beta*C`` in when it stores the tile (``C_ij`` of the root itself
  is computed in place).

All tensors use the column-major convention of the BLAS seam: a ``rows x cols`` column-major matrix is held as a
row-major torch tensor of shape ``(cols, rows)``.
"""
import os

DEFAULT_GRIDS = {1: (1, 1), 2: (1, 2), 4: (1, 4), 8: (1, 8)}


def grid_for(world):
    """pr x pc process grid.  Default 1 x world (column slabs, exactly the reference's split, and the only cut for
    which every panel of a column-major operand is contiguous); ``B200BLAS_GRID=2x4`` selects a 2-D grid."""
    env = os.environ.get("B200BLAS_GRID")
    if env:
        pr, pc = (int(x) for x in env.lower().split("x"))
        if pr * pc != world:
            raise ValueError("B200BLAS_GRID=%s does not match world size %d" % (env, world))
        return pr, pc
    if world in DEFAULT_GRIDS:
        return DEFAULT_GRIDS[world]
    return 1, world


def split(extent, parts, quantum):
    """Cut [0, extent) into `parts` consecutive ranges whose lengths are multiples of `quantum` except the last,
    which takes the remainder -- the rule of Parallelizer.h:140-151 (blockCols & ~3, blockRows rounded to mr) with
    the CTA tile edge as quantum."""
    block = -(-extent // parts)
    block = -(-block // quantum) * quantum
    out = []
    for p in range(parts):
        lo = min(extent, p * block)
        hi = extent if p == parts - 1 else min(extent, (p + 1) * block)
        out.append((lo, max(lo, hi)))
    return out


def partition(m, n, world, grid=None, quantum=128):
    """Tile of C owned by each rank: list of (r0, r1, c0, c1), rank = i*pc + j."""
    pr, pc = grid or grid_for(world)
    rows, cols = split(m, pr, quantum), split(n, pc, quantum)
    return [(rows[i][0], rows[i][1], cols[j][0], cols[j][1]) for i in range(pr) for j in range(pc)]


def chunk_ranges(k, nchunks, quantum=256):
    """Equal k-chunks (nchunks > 0), or -- nchunks == 0 -- the doubling schedule k/16, k/16, k/8, k/4, k/2.
    Compute starts after 1/16 of the panel traffic; each later chunk is as large as everything before it, so it has
    landed by the time it is needed whenever the links deliver panels at least twice as fast as the DMMA pipe
    consumes them (measured at 8 GPUs: 12.5 ms of root egress against 33.5 ms of compute), and only five launches
    pay the per-launch costs (pipeline fill, C read-modify-write, wave tails)."""
    if nchunks == 0:
        if k < 16 * quantum:
            return [(0, k)]
        u = max(quantum, (k // 16) // quantum * quantum)
        cuts = [0, u, 2 * u, 4 * u, 8 * u, k]
        return [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
    return [r for r in split(k, max(1, nchunks), quantum) if r[1] > r[0]]


def subslab_ranges(width, nsub, quantum=128):
    """Column sub-slabs of a tile for the gather pipeline.  nsub == 0: 1/2 + 3/8 + 1/8, so that only an eighth of
    the tile is still on the wire when the last product finishes."""
    if nsub == 0:
        if width < 16 * quantum:
            return [(0, width)] if width > 0 else []
        c1 = (width // 2) // quantum * quantum
        c2 = (width * 7 // 8) // quantum * quantum
        return [(0, c1), (c1, c2), (c2, width)]
    return [r for r in split(width, max(1, nsub), quantum) if r[1] > r[0]]


class DistGemm:
    """C = alpha*A*B + beta*C across all ranks of the default process group (operands root-resident)."""

    def __init__(self, t, m, n, k, alpha, beta, kchunks=0, subslabs=0, grid=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.t, self.m, self.n, self.k, self.alpha, self.beta = t, m, n, k, alpha, beta
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.pr, self.pc = grid or grid_for(self.world)
        self.i, self.j = divmod(self.rank, self.pc)
        self.tiles = partition(m, n, self.world, (self.pr, self.pc))
        self.r0, self.r1, self.c0, self.c1 = self.tiles[self.rank]
        self.mi, self.nj = self.r1 - self.r0, self.c1 - self.c0
        self.chunks = chunk_ranges(k, kchunks)
        self.nsub_arg = subslabs
        self.sub = [(a + self.c0, b + self.c0) for a, b in subslab_ranges(self.nj, subslabs)]
        self.nsub = max(len(subslab_ranges(tl[3] - tl[2], subslabs)) for tl in self.tiles)
        self.dtype = {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128}[t]
        self.backend = dist.get_backend()
        self.dev = torch.device("cuda", torch.cuda.current_device()) if self.backend == "nccl" else torch.device("cpu")
        import eigen_b200
        self.gemm = eigen_b200.gemm_dev  # the sm_100a library; there is no other implementation
        # sub-communicators: grid rows share A_i, grid columns share B_j (every rank creates every group)
        self.row_groups, self.col_groups = [], []
        for i in range(self.pr):
            ranks = [i * self.pc + j for j in range(self.pc)]
            self.row_groups.append(dist.new_group(ranks) if self.pc > 1 and self.pr > 1 else None)
        for j in range(self.pc):
            ranks = [i * self.pc + j for i in range(self.pr)]
            self.col_groups.append(dist.new_group(ranks) if self.pr > 1 and self.pc > 1 else None)
        is_root = self.rank == 0
        kw = dict(dtype=self.dtype, device=self.dev)
        # resident panels (column-major mi x k and k x nj); the root reads its own panels straight from A and B
        self.Ai = None if is_root else torch.empty(k, self.mi, **kw)
        # B_j is assembled as ONE column-major k x nj panel (ld = k) so that phase 2 is a single full-k launch per
        # sub-slab; NCCL needs contiguous buffers, so chunks land in a small staging pair and are copied into place
        self.Bj = None if is_root else torch.empty(self.nj, k, **kw)
        kmax = max(c[1] - c[0] for c in self.chunks)
        self.Bstage = None if is_root else [torch.empty(self.nj * kmax, **kw) for _ in range(2)]
        self.P = None if is_root else torch.empty(self.nj, self.mi, **kw)
        self.recv = None
        if is_root:
            self.recv = {r: torch.empty(tl[3] - tl[2], tl[1] - tl[0], **kw) for r, tl in enumerate(self.tiles) if r != 0}
        if self.dev.type == "cuda":
            self.comm = torch.cuda.Stream(priority=-1)   # panel traffic outranks the GEMM CTAs already queued
            self.out = torch.cuda.Stream(priority=-1)
        else:
            self.comm = self.out = None

    # -- helpers ------------------------------------------------------------------------------------------------
    def _on(self, stream):
        import contextlib
        return self.torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()

    def _event(self, stream=None):
        if self.dev.type != "cuda":
            return None
        e = self.torch.cuda.Event()
        e.record(stream if stream is not None else self.torch.cuda.current_stream())
        return e

    def _wait(self, stream, ev):
        if ev is not None:
            (stream if stream is not None else self.torch.cuda.current_stream()).wait_event(ev)

    def _distribute_chunk(self, A, B, ci):
        """Move chunk ci of every rank's panels (comm stream).  Collective order is identical on all ranks."""
        dist = self.dist
        k0, k1 = self.chunks[ci]
        # ---- A row panels: At[k0:k1, r0:r1] ----
        if self.pr == 1:
            buf = A[k0:k1] if self.rank == 0 else self.Ai[k0:k1]   # contiguous (kc, m): no packing needed
            dist.broadcast(buf, src=0)
        else:
            for i in range(self.pr):
                leader = i * self.pc
                rows = split(self.m, self.pr, 128)[i]
                if self.rank == 0 and leader != 0:
                    dist.send(A[k0:k1, rows[0]:rows[1]].contiguous(), dst=leader)
                elif self.rank == leader and leader != 0:
                    dist.recv(self.Ai[k0:k1], src=0)
                if self.i == i and self.pc > 1:
                    if self.rank == 0:
                        buf = A[k0:k1, rows[0]:rows[1]].contiguous()
                    else:
                        buf = self.Ai[k0:k1]
                    dist.broadcast(buf, src=leader, group=self.row_groups[i])
        # ---- B column panels: Bt[c0:c1, k0:k1]; the root's sends of one chunk go out as ONE grouped NCCL launch ----
        p2p, mine = [], None

        def flush():
            if p2p:
                for req in dist.batch_isend_irecv(p2p):
                    req.wait()
                del p2p[:]

        for j in range(self.pc):
            leader = j
            cols = split(self.n, self.pc, 128)[j]
            stage = None
            if self.rank != 0 and self.j == j:
                stage = self.Bstage[ci % 2][:self.nj * (k1 - k0)].view(self.nj, k1 - k0)
            if self.pc == 1:
                buf = B[:, k0:k1].contiguous() if self.rank == 0 else stage
                dist.broadcast(buf, src=0)
            else:
                if self.rank == 0 and leader != 0:
                    p2p.append(dist.P2POp(dist.isend, B[cols[0]:cols[1], k0:k1].contiguous(), leader))
                elif self.rank == leader and leader != 0:
                    p2p.append(dist.P2POp(dist.irecv, stage, 0))
                if self.pr > 1:
                    flush()
                if self.j == j and self.pr > 1:
                    buf = B[cols[0]:cols[1], k0:k1].contiguous() if self.rank == 0 else stage
                    dist.broadcast(buf, src=leader, group=self.col_groups[j])
            if stage is not None:
                mine = stage
        flush()
        if mine is not None:
            self.Bj[:, k0:k1].copy_(mine)   # strided device copy on the comm stream

    def _local(self, A, B, C, cols, ks, first):
        """One local product on this rank's tile: columns `cols` (global), k range `ks`."""
        c0, c1 = cols
        k0, k1 = ks
        nn, kk = c1 - c0, k1 - k0
        if nn <= 0 or self.mi <= 0:
            return
        if self.rank == 0:
            # operands and the C tile in place inside the caller's matrices (ld = m / k / m)
            a = A[k0:k1, self.r0:]
            b = B[c0:c1, k0:]
            c = C[c0:c1, self.r0:]
            beta = self.beta if first else 1.0
            self.gemm(self.t, "N", "N", self.mi, nn, kk, self.alpha, a, self.m, b, self.k, beta, c, self.m)
        else:
            a = self.Ai[k0:k1]
            b = self.Bj[c0 - self.c0:c1 - self.c0, k0:]
            c = self.P[c0 - self.c0:c1 - self.c0]
            beta = 0.0 if first else 1.0
            self.gemm(self.t, "N", "N", self.mi, nn, kk, self.alpha, a, self.mi, b, self.k, beta, c, self.mi)

    # -- the product ----------------------------------------------------------------------------------------------
    def run(self, A=None, B=None, C=None):
        """A: (k, m), B: (n, k), C: (n, m) torch tensors on rank 0 (None elsewhere).  C is updated in place.

        Schedule: k-chunks of the panels stream over NVLink on the comm stream while the compute stream applies every
        chunk that has landed to the whole local tile (beta only on the first chunk).  The LAST chunk is applied
        sub-slab by sub-slab, and each finished column sub-slab goes back to the root on the out stream while the
        next one computes; the root folds beta*C in as the tiles arrive."""
        torch, dist = self.torch, self.dist
        cur = torch.cuda.current_stream() if self.dev.type == "cuda" else None
        start = self._event(cur)
        self._wait(self.comm, start)
        self._wait(self.out, start)
        ready = []
        for ci in range(len(self.chunks)):
            with self._on(self.comm):
                self._distribute_chunk(A, B, ci)
                ready.append(self._event(self.comm))
        last = len(self.chunks) - 1
        whole = (self.c0, self.c1)
        sends = []
        for ci, ks in enumerate(self.chunks):
            self._wait(cur, ready[ci])
            if not self.sub:
                continue
            if ci < last:
                self._local(A, B, C, whole, ks, ci == 0)
            else:
                for si, cols in enumerate(self.sub):
                    self._local(A, B, C, cols, ks, ci == 0)
                    sends.append((si, self._event(cur)))
        with self._on(self.out):
            for si in range(self.nsub):
                if si < len(sends):
                    self._wait(self.out, sends[si][1])
                self._gather_subslab(C, si)
            fin = self._event(self.out)
        self._wait(cur, fin)
        if self.comm is not None:
            cur.wait_stream(self.comm)

    def _gather_subslab(self, C, si):
        dist = self.dist
        ops, folds = [], []
        if self.rank != 0:
            if si < len(self.sub):
                c0, c1 = self.sub[si]
                ops.append(dist.P2POp(dist.isend, self.P[c0 - self.c0:c1 - self.c0], 0))
        else:
            for r, tl in enumerate(self.tiles):
                if r == 0:
                    continue
                subs = [(a + tl[2], b + tl[2]) for a, b in subslab_ranges(tl[3] - tl[2], self.nsub_arg)]
                if si >= len(subs) or tl[1] <= tl[0]:
                    continue
                s0, s1 = subs[si]
                buf = self.recv[r][s0 - tl[2]:s1 - tl[2]]
                ops.append(dist.P2POp(dist.irecv, buf, r))
                folds.append((buf, C[s0:s1, tl[0]:tl[1]]))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for buf, dst in folds:
            if self.beta == 0:
                dst.copy_(buf)
            else:
                if self.beta != 1:
                    dst.mul_(self.beta)
                dst.add_(buf)

    # -- host-origin product (operands in shared, pinned host memory) ---------------------------------------------
    def run_host(self, hA, hB, hC):
        """End-to-end variant: A (k, m), B (n, k), C (n, m) are CPU tensors that every rank can address (POSIX shared
        memory mapped and cudaHostRegister'ed by each process).  Each GPU pulls only its share over its OWN PCIe link:
        1/N of A (k-block `rank`, then NCCL all-gather over NVLink), its column panel B_j and tile C_j, multiplies,
        and writes C_j straight back into the caller's matrix -- no funnel through GPU 0.  Column-slab grid only."""
        torch, dist = self.torch, self.dist
        assert self.pr == 1, "run_host uses the 1 x N column-slab grid"
        cur = torch.cuda.current_stream()
        kw = dict(dtype=self.dtype, device=self.dev)
        if not hasattr(self, "hA_full"):
            self.hA_full = torch.empty(self.k, self.m, **kw)
            self.hB_j = torch.empty(max(self.nj, 1), self.k, **kw)
            self.hC_j = torch.empty(max(self.nj, 1), self.m, **kw)
        start = self._event(cur)
        self._wait(self.comm, start)
        self._wait(self.out, start)
        even = self.k % self.world == 0
        with self._on(self.comm):
            if even:
                kb = self.k // self.world
                mine = self.hA_full[self.rank * kb:(self.rank + 1) * kb]
                mine.copy_(hA[self.rank * kb:(self.rank + 1) * kb], non_blocking=True)
                dist.all_gather_into_tensor(self.hA_full, mine)
            else:
                self.hA_full.copy_(hA, non_blocking=True)
            a_ready = self._event(self.comm)
        # B_j / C_j sub-slabs are uploaded on the out stream's copy queue in the order they are consumed, so that the
        # upload of sub-slab s+1 overlaps the product on sub-slab s
        hsub = [(a + self.c0, b + self.c0) for a, b in split(self.nj, 4, 128) if b > a]
        up = []
        with self._on(self.out):
            for (s0, s1) in hsub:
                self.hB_j[s0 - self.c0:s1 - self.c0].copy_(hB[s0:s1], non_blocking=True)
                if self.beta != 0:
                    self.hC_j[s0 - self.c0:s1 - self.c0].copy_(hC[s0:s1], non_blocking=True)
                up.append(self._event(self.out))
        self._wait(cur, a_ready)
        evs = []
        for (s0, s1), u in zip(hsub, up):
            self._wait(cur, u)
            self.gemm(self.t, "N", "N", self.m, s1 - s0, self.k, self.alpha, self.hA_full, self.m,
                      self.hB_j[s0 - self.c0:s1 - self.c0], self.k, self.beta, self.hC_j[s0 - self.c0:s1 - self.c0], self.m)
            evs.append(self._event(cur))
        with self._on(self.comm):   # downloads use the other copy queue (D2H engine), behind the A gather
            for (s0, s1), ev in zip(hsub, evs):
                self._wait(self.comm, ev)
                hC[s0:s1].copy_(self.hC_j[s0 - self.c0:s1 - self.c0], non_blocking=True)
            fin = self._event(self.comm)
        self._wait(cur, fin)
        cur.wait_stream(self.comm)
        cur.wait_stream(self.out)
        self.h2d_bytes = (self.k // self.world if even else self.k) * self.m * hA.element_size() + \
            self.nj * self.k * hB.element_size() + (self.nj * self.m * hC.element_size() if self.beta != 0 else 0)
        self.d2h_bytes = self.nj * self.m * hC.element_size()


def shared_host_tensor(name, shape, dtype, create):
    """A CPU tensor backed by POSIX shared memory (/dev/shm/<name>), mapped by every rank and page-locked for CUDA in
    the calling process.  `create` = True on the rank that owns the data."""
    import torch
    numel = 1
    for d in shape:
        numel *= d
    path = "/dev/shm/" + name
    if create and os.path.exists(path):
        os.unlink(path)
    t = torch.from_file(path, shared=True, size=numel, dtype=dtype)
    return t.view(*shape)


def pin_host_range(t):
    """cudaHostRegister the memory of a (contiguous) CPU tensor view in this process."""
    import torch
    rt = torch.cuda.cudart()
    err = rt.cudaHostRegister(t.data_ptr(), t.numel() * t.element_size(), 0)
    if int(err) != 0:
        raise RuntimeError("cudaHostRegister failed: %s" % (err,))
