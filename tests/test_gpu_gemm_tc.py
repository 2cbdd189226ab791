"""The tcgen05/TMEM/TMA GEMM (csrc/gemm_tc.cu) against torch on bf16-rounded operands."""
import pytest
import torch

from ee274_convexcaldera_llm_quantization_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _run(M, N, K, splitk, return_C=False, lda=None, ldb=None, alpha=1.0):
    lib = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(M + 3 * N + 7 * K + splitk)
    lda = lda or (K + 7) // 8 * 8
    ldb = ldb or (K + 7) // 8 * 8
    A = torch.zeros(M, lda, device=DEV, dtype=torch.bfloat16)
    B = torch.zeros(N, ldb, device=DEV, dtype=torch.bfloat16)
    A[:, :K] = torch.randn(M, K, generator=g, device=DEV).bfloat16()
    B[:, :K] = torch.randn(N, K, generator=g, device=DEV).bfloat16()
    if lda > K:
        A[:, K:] = 7.0      # padding beyond K must never be read (TMA tensor map is K wide)
    if ldb > K:
        B[:, K:] = -5.0
    C = torch.full((M, N), float("nan"), device=DEV)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    ws = torch.empty(lib.cb_gemm_bf16_tn_workspace_bytes(), dtype=torch.uint8, device=DEV)
    st = lib.cb_gemm_bf16_tn(M, N, K, alpha, _lib.ptr(A), lda, _lib.ptr(B), ldb, _lib.ptr(C), N, splitk,
                             0, _lib.ptr(flag), _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
    _lib.check(st, "gemm_bf16_tn")
    torch.cuda.synchronize()
    assert int(flag.item()) == 0, "pipeline watchdog fired"
    ref = alpha * (A[:, :K].double() @ B[:, :K].double().T)
    err = float((C.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
    if return_C:
        return err, C
    return err


@pytest.mark.parametrize("M,N,K,splitk", [
    (128, 64, 64, 1), (128, 128, 128, 1), (128, 256, 256, 1), (256, 256, 4096, 1), (256, 256, 4096, 0),
    (256, 4096, 4096, 0), (4096, 256, 4096, 0), (4096, 224, 4096, 0), (224, 224, 4096, 0), (128, 4096, 224, 1),
    (100, 72, 200, 1), (1000, 333, 777, 3), (4096, 4096, 128, 1), (37, 53, 29, 1)])
def test_gemm_tc_matches_torch(M, N, K, splitk):
    err = _run(M, N, K, splitk)
    assert err < 2e-5, err


@pytest.mark.parametrize("M,N,K,splitk", [(224, 224, 4096, 0), (4096, 128, 4096, 0), (1000, 333, 777, 3)])
def test_gemm_tc_splitk_is_bitwise_reproducible(M, N, K, splitk):
    # split-K partial tiles are summed in slice order by the last CTA of each tile, never by
    # floating-point atomics: repeated launches (which also re-use the arrival counters) agree bit for bit
    _, C0 = _run(M, N, K, splitk, return_C=True)
    for _ in range(4):
        _, C1 = _run(M, N, K, splitk, return_C=True)
        assert torch.equal(C0, C1)


def test_gemm_tc_split_without_workspace_is_an_error():
    lib = _lib.load()
    A = torch.zeros(128, 256, device=DEV, dtype=torch.bfloat16)
    C = torch.zeros(128, 128, device=DEV)
    st = lib.cb_gemm_bf16_tn(128, 128, 256, 1.0, _lib.ptr(A), 256, _lib.ptr(A), 256, _lib.ptr(C), 128, 2, 0, None,
                             None, 0, _lib.stream_ptr())
    assert st == -4   # CB_ERR_WORKSPACE


def _tiled(X):
    """(rows, K) -> contiguous 64 x 64 tiles, K blocks of one row block adjacent."""
    r, k = X.shape
    return X.view(r // 64, 64, k // 64, 64).permute(0, 2, 1, 3).contiguous()


@pytest.mark.parametrize("M,N,K,layout", [(256, 4096, 4096, 2), (4096, 128, 4096, 1), (1024, 704, 1536, 3),
                                           (192, 11008, 4096, 2), (4096, 256, 11008, 1)])
def test_gemm_tc_tiled_operands(M, N, K, layout):
    """Operands kept as contiguous 64 x 64 tiles (the driver's HBM layout for the m x n residual):
    bit-identical to the same contraction on row-major operands."""
    lib = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g, device=DEV).bfloat16()
    B = torch.randn(N, K, generator=g, device=DEV).bfloat16()
    ws = torch.empty(lib.cb_gemm_bf16_tn_workspace_bytes(), dtype=torch.uint8, device=DEV)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    out = []
    for lay in (0, layout):
        Aop = _tiled(A) if lay & 1 else A
        Bop = _tiled(B) if lay & 2 else B
        C = torch.full((M, N), float("nan"), device=DEV)
        _lib.check(lib.cb_gemm_bf16_tn(M, N, K, 1.0, _lib.ptr(Aop), K, _lib.ptr(Bop), K, _lib.ptr(C), N, 0, lay,
                                       _lib.ptr(flag), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "gemm")
        out.append(C)
    torch.cuda.synchronize()
    assert int(flag.item()) == 0
    assert torch.equal(out[0], out[1])
    ref = A.double() @ B.double().T
    assert float((out[1].double() - ref).abs().max() / ref.abs().max()) < 2e-5


def test_gemm_tc_alpha_and_ld():
    assert _run(256, 192, 320, 1, lda=384, ldb=512, alpha=-0.5) < 2e-5


def test_convert_bf16():
    lib = _lib.load()
    X = torch.randn(300, 200, device=DEV)
    s = torch.rand(200, device=DEV) + 0.5
    Y = torch.empty(300, 200, device=DEV, dtype=torch.bfloat16)
    Yt = torch.empty(200, 304, device=DEV, dtype=torch.bfloat16)
    _lib.check(lib.cb_convert_bf16(_lib.ptr(X), 300, 200, 200, _lib.ptr(Y), 200, _lib.ptr(Yt), 304, _lib.ptr(s),
                                   _lib.stream_ptr()), "convert")
    ref = (X * s[None, :]).bfloat16()
    assert torch.equal(Y, ref) and torch.equal(Yt[:, :300], ref.T)
