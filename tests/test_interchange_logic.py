"""The bookkeeping behind `perm_apply_direct_kernel` and the chained LU leaf (eigen_b200/csrc/lapack.cu), restated in numpy and
checked against the definition: xLASWP applies the interchanges k <-> piv_k for k ascending (PartialPivLU.h:384-388,
lapack/lu.cpp).  The kernels never run that chain of dependent swaps; they note that only the rows {k} and {piv_k} can
change, find the SOURCE row of each such destination by undoing the interchanges in reverse order, read every source, then
write every destination.  CPU-only: this pins the algorithm, the GPU tests pin the kernels."""
import numpy as np
import pytest


def laswp(a, piv):
    out = a.copy()
    for k, p in enumerate(piv):
        if p != k:
            out[[k, p]] = out[[p, k]]
    return out


def source_row(pos, piv, k_first=0):
    for k in range(len(piv) - 1, -1, -1):
        kk, pk = k_first + k, piv[k]
        if pos == kk:
            pos = pk
        elif pos == pk:
            pos = kk
    return pos


def direct_apply(a, piv):
    ns = len(piv)
    dsts = list(range(ns)) + [p for p in piv if p >= ns]          # threads 0..ns-1 and ns..2ns-1 of the kernel
    srcs = [source_row(d, piv) for d in dsts]
    vals = [a[s].copy() for s in srcs]                             # read every source ...
    out = a.copy()
    for d, v in zip(dsts, vals):                                   # ... then write every destination (duplicates carry the same value)
        out[d] = v
    return out


@pytest.mark.parametrize("seed", range(20))
def test_direct_interchange_equals_sequential_swaps(seed):
    rng = np.random.default_rng(seed)
    m = int(rng.integers(1, 400))
    ns = int(rng.integers(1, min(m, 128) + 1))
    # pivots as partial pivoting produces them: piv_k >= k; include repeats and piv_k == k
    piv = [int(rng.integers(k, m)) if rng.random() < 0.8 else k for k in range(ns)]
    if ns > 2 and m > ns:
        piv[1] = piv[0]                                            # the same far row hit twice
    a = rng.standard_normal((m, 5))
    assert np.array_equal(direct_apply(a, piv), laswp(a, piv))


@pytest.mark.parametrize("seed", range(10))
def test_gather_for_the_second_phase_of_the_chained_leaf(seed):
    """Phase 1 of getf2_reg2_kernel loads row `source_row(r)` of the untouched columns instead of swapping them 32 times."""
    rng = np.random.default_rng(100 + seed)
    m, nbp = int(rng.integers(64, 300)), 32
    piv = [int(rng.integers(k, m)) for k in range(nbp)]
    a = rng.standard_normal((m, 7))
    gathered = np.stack([a[source_row(r, piv)] for r in range(m)])
    assert np.array_equal(gathered, laswp(a, piv))
    # and the final fix-up: phase 1's interchanges (rows nbp ..) applied to the L part of phase 0's columns
    piv1 = [int(rng.integers(nbp + k, m)) for k in range(nbp)]
    want = a.copy()
    for k, p in enumerate(piv1):
        if p != nbp + k:
            want[[nbp + k, p]] = want[[p, nbp + k]]
    dsts = [nbp + t for t in range(nbp)] + [p for p in piv1 if p >= 2 * nbp]
    got = a.copy()
    vals = [a[source_row(d, piv1, k_first=nbp)].copy() for d in dsts]
    for d, v in zip(dsts, vals):
        got[d] = v
    assert np.array_equal(got, want)
