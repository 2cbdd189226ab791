"""GPU checks of the contraction / factorisation kernels against torch fp64 references:
strided SGEMM (csrc/sgemm.cu), blocked Cholesky + inverse and the one-sided Jacobi
eigensolver (csrc/smalldense.cu), and the randomized rank-r factorisation built from them
(cb_lowrank_init, replaces LR_init, alg.py:201-235)."""
import numpy as np
import pytest
import torch

from ee274_convexcaldera_llm_quantization_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _sgemm(A, B, transA=False, transB=False, accumulate=False, C=None):
    lib = _lib.load()
    Aop = A.T if transA else A
    Bop = B.T if transB else B
    M, K = Aop.shape
    N = Bop.shape[1]
    if C is None:
        C = torch.empty(M, N, device=DEV)
    st = lib.cb_sgemm_strided(M, N, K, 1.0, _lib.ptr(A), Aop.stride(0), Aop.stride(1), _lib.ptr(B),
                              Bop.stride(0), Bop.stride(1), _lib.ptr(C), C.stride(0), C.stride(1),
                              int(accumulate), _lib.stream_ptr())
    _lib.check(st, "sgemm")
    return C


@pytest.mark.parametrize("M,N,K", [(37, 53, 29), (64, 64, 4096), (256, 256, 4096), (512, 640, 300),
                                   (4096, 256, 512), (1, 1, 1), (130, 7, 1000), (1024, 1024, 128)])
@pytest.mark.parametrize("tA,tB", [(False, False), (True, False), (False, True), (True, True)])
def test_sgemm_strided(M, N, K, tA, tB):
    g = torch.Generator(device=DEV).manual_seed(M * 7 + N * 3 + K)
    A = torch.randn((K, M) if tA else (M, K), generator=g, device=DEV)
    B = torch.randn((N, K) if tB else (K, N), generator=g, device=DEV)
    C = _sgemm(A, B, tA, tB)
    ref = ((A.T if tA else A).double() @ (B.T if tB else B).double())
    err = float((C.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
    assert err < 2e-5, err
    C2 = _sgemm(A, B, tA, tB, accumulate=True, C=C.clone())
    err2 = float((C2.double() - 2 * ref).abs().max() / ref.abs().max().clamp_min(1e-30))
    assert err2 < 4e-5, err2


@pytest.mark.parametrize("q", [1, 8, 31, 32, 33, 40, 96, 128, 256, 300, 512])
def test_cholesky_inverse(q):
    lib = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(q)
    A = torch.randn(q, q + 16, generator=g, device=DEV, dtype=torch.float64)
    G64 = A @ A.T / (q + 16) + 0.05 * torch.eye(q, device=DEV, dtype=torch.float64)
    G = G64.float().contiguous()
    G0 = G.clone()
    Linv = torch.full((q, q), float("nan"), device=DEV)
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    _lib.check(lib.cb_cholesky_inverse_f32(_lib.ptr(G), q, _lib.ptr(Linv), _lib.ptr(status), _lib.stream_ptr()), "chol")
    assert int(status.item()) == 0
    Lc = torch.tril(G).double()
    ref = torch.linalg.cholesky(G0.double())
    assert float((Lc - ref).abs().max() / ref.abs().max()) < 5e-5
    # the strict upper triangle is left untouched (it is the retry backup)
    assert torch.equal(torch.triu(G, 1), torch.triu(G0, 1))
    eye = Linv.double() @ ref
    assert torch.isfinite(Linv).all()
    assert float((eye - torch.eye(q, device=DEV, dtype=torch.float64)).abs().max()) < 2e-3
    assert float(torch.triu(Linv, 1).abs().max()) == 0.0 if q > 1 else True


def test_cholesky_ridge_retry():
    lib = _lib.load()
    q = 64
    v = torch.randn(q, 8, device=DEV)
    G = (v @ v.T).contiguous()          # rank 8: plain Cholesky must break down
    Linv = torch.empty(q, q, device=DEV)
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    _lib.check(lib.cb_cholesky_inverse_f32(_lib.ptr(G), q, _lib.ptr(Linv), _lib.ptr(status), _lib.stream_ptr()), "chol")
    assert 1 <= int(status.item()) <= 3
    assert torch.isfinite(Linv).all() and torch.isfinite(torch.tril(G)).all()


@pytest.mark.parametrize("q", [2, 7, 32, 64, 100, 256, 512])
def test_jacobi_eigh(q):
    lib = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(q + 1)
    A = torch.randn(q, 3 * q, generator=g, device=DEV, dtype=torch.float64)
    A = A * torch.linspace(1.0, 0.2, q, device=DEV, dtype=torch.float64)[:, None]
    G64 = A @ A.T / (3 * q)
    Lc = torch.linalg.cholesky(G64).float().contiguous()
    Lc = Lc + torch.triu(torch.full((q, q), 123.0, device=DEV), 1)   # upper triangle must be ignored
    evals = torch.empty(q, device=DEV)
    evecs = torch.empty(q, q, device=DEV)
    work = torch.empty(q * q + q + 8, device=DEV)
    sweeps = torch.zeros(1, dtype=torch.int32, device=DEV)
    _lib.check(lib.cb_jacobi_eigh_from_chol_f32(_lib.ptr(Lc), q, _lib.ptr(evals), _lib.ptr(evecs), _lib.ptr(work),
                                                _lib.ptr(sweeps), _lib.stream_ptr()), "jacobi")
    ref = torch.linalg.eigvalsh(G64).flip(0)
    assert 1 <= int(sweeps.item()) <= 30
    # ~q rotations per column and sweep, each norm-preserving to an ulp: allow 1e-4 relative drift
    assert float((evals.double() - ref).abs().max() / ref.max()) < 2e-4
    V = evecs.double()                          # rows are eigenvectors
    assert float((V @ V.T - torch.eye(q, device=DEV, dtype=torch.float64)).abs().max()) < 2e-4   # kJacobiTol = 3e-5
    resid = G64 @ V.T - V.T * evals.double()[None, :]
    assert float(resid.abs().max() / ref.max()) < 2e-4
    assert bool((evals[:-1] >= evals[1:]).all())


@pytest.mark.parametrize("m,n,r,aware", [(512, 384, 32, 1), (384, 512, 16, 0), (300, 200, 10, 1)])
def test_lowrank_init_vs_svd(m, n, r, aware):
    lib = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(m + n)
    U = torch.linalg.qr(torch.randn(m, m, generator=g, device=DEV))[0]
    Vt = torch.linalg.qr(torch.randn(n, n, generator=g, device=DEV))[0]
    k = min(m, n)
    s = torch.arange(1, k + 1, device=DEV).float() ** -0.6
    A = (U[:, :k] * s) @ Vt[:k, :]
    h = 0.5 + torch.rand(n, generator=g, device=DEV)
    q = min(max(2 * r, r + 32), k)
    L = torch.empty(m, r, device=DEV)
    R = torch.empty(r, n, device=DEV)
    sig = torch.empty(r, device=DEV)
    ws_bytes = lib.cb_lowrank_init_workspace_bytes(m, n, r, q, _lib.CB_H_DIAG)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=DEV)
    _lib.check(lib.cb_lowrank_init(_lib.ptr(A), m, n, _lib.ptr(h), _lib.CB_H_DIAG, r, q, 6, 1234, aware,
                                   _lib.ptr(L), _lib.ptr(R), _lib.ptr(sig), _lib.ptr(ws), ws_bytes,
                                   _lib.stream_ptr()), "lowrank_init")
    Y = (A * h.sqrt()[None, :]) if aware else A
    S = torch.linalg.svdvals(Y.double())
    opt = float((S[r:] ** 2).sum().sqrt())
    E = (A - L @ R).double()
    got = float(((E * h.sqrt()[None, :].double()) if aware else E).norm())
    assert got <= opt * (1 + 2e-4), (got, opt)
    assert float((sig.double() - S[:r]).abs().max() / S[0]) < 1e-3
    if aware:
        assert float((L.T @ L - torch.eye(r, device=DEV)).abs().max()) < 1e-4
