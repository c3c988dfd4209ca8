"""The multi-GPU partition plan (include/b200blas.h section 3, eigen_b200/csrc/multi.cu) checked on the CPU.

The library builds the partition of one product as DATA -- copy / product / axpby steps with device, stream slot and
dependencies -- and its executor replays that list on CUDA streams.  Here the same list is
  1. interpreted with numpy (device buffers start as NaN, so reading anything that was never delivered poisons C) on
     small-integer operands, for which every summation order gives the exact result: C must EQUAL alpha*op(A)*op(B)+beta*C;
  2. race-checked: every two steps that touch overlapping regions of one buffer, at least one of them writing, must be
     ordered by stream order + wait edges (the executor runs streams concurrently);
  3. measured: every element of A, B and C crosses the slow link (the origin) exactly once.
Reference role: parallelize_gemm's partition (Eigen/src/Core/products/Parallelizer.h:140-154) -- every element of C is
owned by exactly one worker and k is never split.  No GPU is needed: the plan is pure host arithmetic.
"""
import itertools

import numpy as np
import pytest

import eigen_b200

COPY, GEMM, AXPBY = 0, 1, 2
NP_DT = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}


def _op(x, o):
    return x if o == 0 else x.T if o == 1 else x.conj().T


def _int_matrix(rng, t, rows, cols):
    x = rng.integers(-2, 3, size=(rows, cols)).astype(np.float64)
    if t in "cz":
        x = x + 1j * rng.integers(-2, 3, size=(rows, cols))
    return np.asfortranarray(x.astype(NP_DT[t]))


class Interp:
    def __init__(self, t, info, A, B, C):
        self.t, self.info = t, info
        self.buf = {(-1, 0): A, (-1, 1): B, (-1, 2): C}
        for d in range(info.ndev):
            for b in range(4):
                if info.elems[d][b] > 0:
                    ld = info.ld[d][b]
                    self.buf[(d, b)] = np.full((ld, info.elems[d][b] // ld), np.nan, dtype=NP_DT[t], order="F")

    def scalar(self, v):
        return complex(v[0], v[1]) if self.t in "cz" else float(v[0])

    def view(self, r):
        key = (r.loc, r.buf)
        if key not in self.buf:
            assert r.loc == 0 and r.buf >= 4, key          # staging tile on the root for the tile of device buf - 4
            src = r.buf - 4
            ld = self.info.ld[src][2]
            self.buf[key] = np.full((ld, self.info.elems[src][2] // ld), np.nan, dtype=NP_DT[self.t], order="F")
        a = self.buf[key]
        assert 0 <= r.r0 and r.r0 + r.rows <= a.shape[0] and 0 <= r.c0 and r.c0 + r.cols <= a.shape[1], (key, r.r0, r.c0, r.rows, r.cols, a.shape)
        return a[r.r0:r.r0 + r.rows, r.c0:r.c0 + r.cols]

    def run(self, steps):
        for s in steps:
            if s.kind == COPY:
                assert (s.x.rows, s.x.cols) == (s.z.rows, s.z.cols)
                self.view(s.z)[...] = self.view(s.x)
            elif s.kind == GEMM:
                al, be = self.scalar(s.alpha), self.scalar(s.beta)
                z = self.view(s.z)
                prod = _op(self.view(s.x), s.opa) @ _op(self.view(s.y), s.opb)
                assert prod.shape == z.shape
                z[...] = (al * prod + (be * z if be != 0 else 0)).astype(z.dtype)    # beta == 0: C is not read
            else:
                al, be = self.scalar(s.alpha), self.scalar(s.beta)
                z = self.view(s.z)
                z[...] = (be * z + al * self.view(s.x)).astype(z.dtype)


def _accesses(s):
    """[(loc, buf, r0, c0, rows, cols, writes)]"""
    def acc(r, w):
        return (r.loc, r.buf, r.r0, r.c0, r.rows, r.cols, w)
    if s.kind == COPY:
        return [acc(s.x, False), acc(s.z, True)]
    if s.kind == GEMM:
        return [acc(s.x, False), acc(s.y, False), acc(s.z, True)]
    return [acc(s.x, False), acc(s.z, True)]


def _overlap(a, b):
    return a[0] == b[0] and a[1] == b[1] and a[2] < b[2] + b[4] and b[2] < a[2] + a[4] and a[3] < b[3] + b[5] and b[3] < a[3] + a[5]


def check_structure_and_races(steps):
    n = len(steps)
    anc = [None] * n           # happens-before ancestors as Python ints used as bitsets
    last = {}
    for i, s in enumerate(steps):
        bits = 0
        prev = last.get((s.dev, s.stream))
        if prev is not None:
            bits |= anc[prev] | (1 << prev)
        assert 0 <= s.nwait <= 4
        for w in list(s.wait)[:s.nwait]:
            assert 0 <= w < i, "a step waits on a later (or missing) step"
            assert steps[w].record == 1, "a waited-on step does not record an event"
            assert (steps[w].dev, steps[w].stream) != (s.dev, s.stream)
            bits |= anc[w] | (1 << w)
        anc[i] = bits
        last[(s.dev, s.stream)] = i
    acc = [_accesses(s) for s in steps]
    for j in range(n):
        for i in range(j):
            if (anc[j] >> i) & 1:
                continue
            for a in acc[i]:
                for b in acc[j]:
                    if (a[6] or b[6]) and _overlap(a, b):
                        raise AssertionError("unordered conflicting steps %d and %d: %r / %r" % (i, j, a, b))


def run_case(t, ta, tb, m, n, k, alpha, beta, ndev, grid, host_origin, seed=0):
    rng = np.random.default_rng(seed)
    info, steps = eigen_b200.multi_plan(t, ta, tb, m, n, k, alpha, beta, ndev, grid=grid, host_origin=host_origin)
    assert info.ndev == ndev and info.pr * info.pc == ndev and len(steps) == info.nsteps
    ra, ca = (m, k) if ta == "N" else (k, m)
    rb, cb = (k, n) if tb == "N" else (n, k)
    A, B, C0 = _int_matrix(rng, t, ra + 3, ca), _int_matrix(rng, t, rb + 1, cb), _int_matrix(rng, t, m + 2, n)   # ld > rows
    if beta == 0:
        C0[:m] = np.nan
    C = C0.copy(order="F")
    o = {"N": 0, "T": 1, "C": 2}
    want = alpha * (_op(A[:ra], o[ta]) @ _op(B[:rb], o[tb])) + (beta * C0[:m] if beta != 0 else 0)
    Interp(t, info, A, B, C).run(steps)
    assert np.array_equal(C[:m], want.astype(NP_DT[t])), "the plan does not compute alpha*op(A)*op(B) + beta*C"
    assert C[m:].tobytes() == C0[m:].tobytes(), "rows beyond m were written"
    check_structure_and_races(steps)
    # every element of the operands crosses the slow link (origin -> a device) exactly once; in the device-resident mode
    # the root reads its own share in place
    seen = {0: np.zeros((ra, ca), np.int32), 1: np.zeros((rb, cb), np.int32)}
    for s in steps:
        if s.kind == COPY and s.x.loc == -1 and s.x.buf in (0, 1) and s.stream == 0:
            seen[s.x.buf][s.x.r0:s.x.r0 + s.x.rows, s.x.c0:s.x.c0 + s.x.cols] += 1
    assert seen[0].max() <= 1 and seen[1].max() <= 1, "an operand element is fetched from the origin twice"
    if host_origin:
        assert seen[0].min() == 1 and seen[1].min() == 1
    return info, steps


@pytest.mark.parametrize("host_origin", [False, True])
@pytest.mark.parametrize("ndev,grid", [(1, (0, 0)), (2, (0, 0)), (4, (0, 0)), (8, (0, 0)), (8, (1, 8)), (8, (4, 2)), (4, (4, 1)), (6, (0, 0))])
def test_plan_computes_the_product_and_is_race_free(ndev, grid, host_origin):
    ops = list(itertools.product("NTC", "NTC"))
    for ci, (m, n, k) in enumerate([(700, 900, 1300), (1500, 2300, 3000), (257, 5000, 600), (3000, 300, 2049)]):
        ta, tb = ops[(ci * 4 + ndev) % 9]
        t = "z" if ci % 2 else "d"
        alpha = (2 - 1j) if t == "z" else 2.0
        beta = [(-3 + 2j) if t == "z" else -3.0, 0.0, 1.0][(ci + ndev) % 3]
        run_case(t, ta, tb, m, n, k, alpha, beta, ndev, grid, host_origin, seed=ci)


def test_default_grids_and_headline_plan_shape():
    """2 -> 1x2, 4 -> 2x2, 8 -> 2x4, tile edges on multiples of the 256-wide pair tile; the 16384^3 plan at 8
    GPUs moves each operand byte over the slow link once and returns C in sub-slabs of 1/2, 3/8, 1/8."""
    for ndev, want in [(1, (1, 1)), (2, (1, 2)), (4, (2, 2)), (8, (2, 4))]:
        info, _ = eigen_b200.multi_plan("d", "N", "N", 16384, 16384, 16384, 1.0, 1.0, ndev, cap=0 + 4096)
        assert (info.pr, info.pc) == want
        rows, cols = list(info.row_cut)[:info.pr + 1], list(info.col_cut)[:info.pc + 1]
        assert rows[0] == 0 and rows[-1] == 16384 and cols[0] == 0 and cols[-1] == 16384
        assert all(c % 256 == 0 for c in rows + cols)
    info, steps = eigen_b200.multi_plan("d", "N", "N", 16384, 16384, 16384, 1.0, 1.0, 8)
    assert info.nchunks == 16 and list(info.group_first_chunk)[:info.ngroups + 1] == [0, 1, 2, 4, 8, 16]
    root_egress = sum(s.x.rows * s.x.cols for s in steps if s.kind == COPY and s.x.loc == -1) * 8
    # all of A and B minus the root's own tile share, plus the chunks the root owns relayed to its row / column mates
    assert root_egress <= 4.6 * 2 ** 30
    ret = [s for s in steps if s.kind == COPY and s.stream == 3 and s.dev == 5]
    assert [r.x.cols for r in ret] == [2048, 1536, 512]
    info, steps = eigen_b200.multi_plan("d", "N", "N", 16384, 16384, 16384, 1.0, 1.0, 8, host_origin=True)
    h2d = sum(s.x.rows * s.x.cols for s in steps if s.kind == COPY and s.x.loc == -1) * 8
    d2h = sum(s.x.rows * s.x.cols for s in steps if s.kind == COPY and s.z.loc == -1) * 8
    assert h2d == 3 * 2 ** 31 and d2h == 2 ** 31     # A, B, C once up; C once down


def test_ragged_and_degenerate_partitions():
    """More devices than 256-wide tiles (some devices get nothing), one chunk, tiny k."""
    run_case("d", "N", "T", 300, 200, 100, 1.0, 1.0, 8, (0, 0), False)
    run_case("d", "T", "N", 300, 200, 100, 1.0, 0.0, 8, (0, 0), True)
    run_case("z", "C", "C", 520, 260, 513, 1.0, 2.0, 4, (0, 0), False)
    run_case("d", "N", "N", 1, 1, 1, 3.0, 2.0, 2, (0, 0), True)
    with pytest.raises(ValueError):
        eigen_b200.multi_plan("d", "N", "N", 100, 100, 100, 1.0, 1.0, 8, grid=(3, 2))
