"""BASELINE.json's configurations at FULL size on the B200, checked through size-independent properties and sampled
rows of the long-double oracle (a full reference of a 16384^3 product would take hours on the CPU).

  C2  dgemm 16384^3            exact integer known answer + 16 sampled rows vs oracle_hp_gemm
  C3  sgemm 8192^3 (3xTF32)    sampled rows, relative Frobenius error <= k*eps, gauge ratio < 16
  C4  zgemm / cgemm 4096^3     sampled rows, xBLAT alpha/beta (0.7,-0.9)/(1.3,-1.1)
  C5  rank-k update            A22 -= A21*A12 on sub-blocks of one 16640^2 matrix (lda = ldb = ldc = 16640)
"""
import numpy as np
import pytest

import eigen_b200
import oracle_api as oa

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

TOL_FRO = 1.0     # relative Frobenius error <= TOL_FRO * k * eps   (BASELINE.json: c*k*eps, c = 1)
TOL_RATIO = 16.0  # netlib gauge ratio, blas/testing/dblat3.dat:8


def _tdt(t):
    return {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128}[t]


def _urand(t, rows, cols, gen):
    """column-major rows x cols on the GPU == row-major (cols, rows) tensor, uniform[-1,1] (bench_gemm.cpp:211-214)"""
    if t in "cz":
        x = torch.rand(cols, rows, 2, dtype=torch.float64 if t == "z" else torch.float32, device="cuda", generator=gen)
        return torch.view_as_complex(x * 2 - 1)
    return torch.rand(cols, rows, dtype=_tdt(t), device="cuda", generator=gen) * 2 - 1


def _sampled_check(t, m, n, k, alpha, beta, A, B, C0, C, rows):
    """Compare the sampled rows of C (device tensors, column-major) with the long-double oracle."""
    An = np.asfortranarray(A.cpu().numpy().T)      # m x k
    Bn = np.asfortranarray(B.cpu().numpy().T)      # k x n
    C0n = np.asfortranarray(C0.cpu().numpy().T)    # m x n
    ref, g = oa.hp_gemm(t, "N", "N", m, n, k, alpha, An, m, Bn, k, beta, C0n, m, rows=rows)
    got = C[:, rows].cpu().numpy().T               # len(rows) x n
    eps = oa.EPS[t]
    ratio = float((np.abs(got - ref) / (eps * g)).max())
    fro = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
    assert ratio < TOL_RATIO, (t, m, n, k, "gauge ratio", ratio)
    assert fro <= TOL_FRO * k * eps, (t, m, n, k, "rel fro", fro, "bound", k * eps)
    return fro, ratio


def test_c2_dgemm_16384_exact_known_answer():
    """Small-integer operands make every partial sum exactly representable, so the 16384^3 product must be EXACT:
    A(i,p) = ((i + 3p) mod 5) - 2, B(p,j) = ((p + 2j) mod 7) - 3, C = 1, C += A*B."""
    n = 16384
    i = torch.arange(n, device="cuda", dtype=torch.int64)
    A = (((i[None, :] + 3 * i[:, None]) % 5) - 2).to(torch.float64)   # (k, m): element (p, i)
    B = (((i[None, :] + 2 * i[:, None]) % 7) - 3).to(torch.float64)   # (n, k): element (j, p)
    C = torch.ones(n, n, dtype=torch.float64, device="cuda")
    assert eigen_b200.gemm_dev("d", "N", "N", n, n, n, 1.0, A, n, B, n, 1.0, C, n) == 0
    torch.cuda.synchronize()
    assert "dmma" in eigen_b200.last_variant()
    rows = np.array([0, 1, 127, 128, 4095, 8191, 8192, 12345, 16383])
    Ar = A[:, rows].cpu().numpy().T                                    # len(rows) x k, exact small integers
    want = Ar @ B.cpu().numpy().T + 1.0                                # exact in float64 (|sum| < 2^53)
    got = C[:, rows].cpu().numpy().T
    assert np.array_equal(got, want)
    # checksum of checksums over the WHOLE result: sum_ij C = sum_p (sum_i A_ip)(sum_j B_pj) + n^2, exact integers
    colsum_a = A.sum(dim=1)
    rowsum_b = B.sum(dim=0)
    assert float(C.sum().item()) == float((colsum_a * rowsum_b).sum().item() + n * n)


def test_c2_dgemm_16384_sampled_rows_vs_oracle():
    n = 16384
    gen = torch.Generator(device="cuda").manual_seed(42)
    A, B = _urand("d", n, n, gen), _urand("d", n, n, gen)
    C0 = torch.ones(n, n, dtype=torch.float64, device="cuda")
    C = C0.clone()
    assert eigen_b200.gemm_dev("d", "N", "N", n, n, n, 1.0, A, n, B, n, 1.0, C, n) == 0
    torch.cuda.synchronize()
    rows = np.array([0, 63, 64, 2047, 5000, 8191, 8192, 16383], dtype=np.int32)
    fro, ratio = _sampled_check("d", n, n, n, 1.0, 1.0, A, B, C0, C, rows)
    print("dgemm 16384^3: rel fro %.3e (k*eps = %.3e), gauge ratio %.3f" % (fro, n * oa.EPS["d"], ratio))


def test_c3_sgemm_8192_3xtf32_sampled_rows_vs_oracle():
    n = 8192
    gen = torch.Generator(device="cuda").manual_seed(43)
    A, B = _urand("s", n, n, gen), _urand("s", n, n, gen)
    C0 = torch.ones(n, n, dtype=torch.float32, device="cuda")
    C = C0.clone()
    assert eigen_b200.gemm_dev("s", "N", "N", n, n, n, 1.0, A, n, B, n, 1.0, C, n) == 0
    torch.cuda.synchronize()
    assert "tf32x3" in eigen_b200.last_variant()
    rows = np.arange(0, n, 257, dtype=np.int32)
    fro, ratio = _sampled_check("s", n, n, n, 1.0, 1.0, A, B, C0, C, rows)
    print("sgemm 8192^3 (3xTF32): rel fro %.3e (k*eps = %.3e), gauge ratio %.3f" % (fro, n * oa.EPS["s"], ratio))
    # idempotence of the data path: the same call on the same inputs is bit-identical
    C2 = C0.clone()
    eigen_b200.gemm_dev("s", "N", "N", n, n, n, 1.0, A, n, B, n, 1.0, C2, n)
    torch.cuda.synchronize()
    assert torch.equal(C, C2)


@pytest.mark.parametrize("t", ["z", "c"])
def test_c4_complex_4096_sampled_rows_vs_oracle(t):
    n = 4096
    gen = torch.Generator(device="cuda").manual_seed(44)
    A, B = _urand(t, n, n, gen), _urand(t, n, n, gen)
    C0 = _urand(t, n, n, gen)
    al, be = 0.7 - 0.9j, 1.3 - 1.1j   # zblat3.dat:12-14
    C = C0.clone()
    assert eigen_b200.gemm_dev(t, "N", "N", n, n, n, al, A, n, B, n, be, C, n) == 0
    torch.cuda.synchronize()
    rows = np.arange(5, n, 521, dtype=np.int32)
    fro, ratio = _sampled_check(t, n, n, n, al, be, A, B, C0, C, rows)
    print("%sgemm 4096^3: rel fro %.3e, gauge ratio %.3f, variant %s" % (t, fro, ratio, eigen_b200.last_variant()))
    # adjoint identity at full size: (A^H B)^H == B^H A, i.e. op pairs C,N and C,N swapped agree
    P = torch.zeros(n, n, dtype=_tdt(t), device="cuda")
    Q = torch.zeros(n, n, dtype=_tdt(t), device="cuda")
    eigen_b200.gemm_dev(t, "C", "N", n, n, n, 1.0, A, n, B, n, 0.0, P, n)   # P = A^H B
    eigen_b200.gemm_dev(t, "C", "N", n, n, n, 1.0, B, n, A, n, 0.0, Q, n)   # Q = B^H A = P^H
    torch.cuda.synchronize()
    # column-major P is the row-major tensor P^T; (P^T)^H-ish comparison: Q == conj(P^T) elementwise on the tensors
    diff = (Q - P.transpose(0, 1).conj()).abs().max().item()
    scale = P.abs().max().item()
    assert diff <= 64 * n * oa.EPS[t] * scale / n ** 0.5


def test_c5_rank_k_trailing_update_in_place_subblocks():
    """LU/PartialPivLU.h:492: A22.noalias() -= A21 * A12 with every operand a sub-block of ONE matrix."""
    N, bs = 16640, 256
    m = N - bs
    gen = torch.Generator(device="cuda").manual_seed(45)
    M = _urand("d", N, N, gen)                     # tensor (col, row); element (i, j) of the matrix at M[j, i]
    M0 = M.clone()
    A21 = M[0:bs, bs:]                             # rows bs.., cols 0..bs   -> pointer of element (bs, 0)
    A12 = M[bs:, 0:bs]                             # rows 0..bs, cols bs..   -> pointer of element (0, bs)
    A22 = M[bs:, bs:]
    assert eigen_b200.gemm_dev("d", "N", "N", m, m, bs, -1.0, A21, N, A12, N, 1.0, A22, N) == 0
    torch.cuda.synchronize()
    assert torch.equal(M[0:bs, :], M0[0:bs, :]) and torch.equal(M[:, 0:bs], M0[:, 0:bs]), "inputs / border were modified"
    rows = np.array([0, 1, 4097, m - 1], dtype=np.int32)
    Mh = np.asfortranarray(M0.cpu().numpy().T)     # N x N column-major
    ref, g = oa.hp_gemm("d", "N", "N", m, m, bs, -1.0, Mh[bs:, :bs], N, Mh[:bs, bs:], N, 1.0, Mh[bs:, bs:], N, rows=rows)
    got = M[bs:, bs:][:, rows].cpu().numpy().T
    assert float((np.abs(got - ref) / (oa.EPS["d"] * g)).max()) < TOL_RATIO
