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


@pytest.mark.parametrize("M,N,K,splitk", [(256, 4096, 4096, 1), (224, 224, 4096, 0), (4096, 128, 4096, 0), (128, 512, 192, 1),
                                           (256, 1024, 4096 + 64, 3)])
def test_gemm_tc_kblocks_per_copy_do_not_change_results(M, N, K, splitk):
    """One or two 64-wide K blocks per TMA instruction (3-D tensor map): same bits.  (Needs the measurement build,
    CB_LIBRARY=libcaldera_b200_measure.so: the release library has no such switch.)"""
    lib = _lib.load()
    if not hasattr(lib, "cb_set_gemm_kblocks"):
        pytest.skip("measurement aids are not compiled into the release library")
    try:
        lib.cb_set_gemm_kblocks(1)
        _, C1 = _run(M, N, K, splitk, return_C=True)
        lib.cb_set_gemm_kblocks(2)
        _, C2 = _run(M, N, K, splitk, return_C=True)
    finally:
        lib.cb_set_gemm_kblocks(2)
    assert torch.equal(C1, C2)


@pytest.mark.parametrize("M,N,K,mode", [(224, 4096, 4096, 0), (224, 4096, 4096, 1), (4096, 224, 224, 0),
                                         (224, 1024, 256, 1), (100, 72, 200, 0), (256, 520, 128, 1)])
@pytest.mark.parametrize("outs", ["both", "rowmajor", "transposed"])
def test_gemm_tc_bf16_epilogues(M, N, K, mode, outs):
    """bf16 row-major / transposed outputs with scaling against torch, for both grid policies (exec_mode of the call:
    wide tiles on ~32 CTAs or narrow tiles over the machine) and for shapes that cannot be staged (ragged N / M);
    with the measurement build also: shared-memory staged epilogue == direct stores, bit for bit."""
    lib = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    Kp = (K + 7) // 8 * 8
    A = torch.zeros(M, Kp, device=DEV, dtype=torch.bfloat16)
    B = torch.zeros(N, Kp, device=DEV, dtype=torch.bfloat16)
    A[:, :K] = torch.randn(M, K, generator=g, device=DEV).bfloat16()
    B[:, :K] = torch.randn(N, K, generator=g, device=DEV).bfloat16()
    cs = torch.rand(N, generator=g, device=DEV) + 0.5
    rs = torch.rand(M, generator=g, device=DEV) + 0.5
    ldcb, ldct = (N + 7) // 8 * 8, (M + 7) // 8 * 8
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    res = []
    have_aids = hasattr(lib, "cb_set_gemm_staged_epilogue")
    try:
        for staged in ((1, 0) if have_aids else (1,)):
            if have_aids:
                lib.cb_set_gemm_staged_epilogue(staged)
            Cb = torch.full((M, ldcb), 7.0, device=DEV, dtype=torch.bfloat16)
            Ct = torch.full((N, ldct), 7.0, device=DEV, dtype=torch.bfloat16)
            _lib.check(lib.cb_gemm_bf16_tn_bf16out(M, N, K, 0.5, _lib.ptr(A), Kp, _lib.ptr(B), Kp,
                                                   _lib.ptr(Cb) if outs != "transposed" else None, ldcb,
                                                   _lib.ptr(Ct) if outs != "rowmajor" else None, ldct,
                                                   _lib.ptr(cs), _lib.ptr(rs), mode, _lib.ptr(flag), _lib.stream_ptr()), "gemm")
            res.append((Cb, Ct))
    finally:
        if have_aids:
            lib.cb_set_gemm_staged_epilogue(1)
    torch.cuda.synchronize()
    assert int(flag.item()) == 0
    ref = (0.5 * (A[:, :K].double() @ B[:, :K].double().T) * rs[:, None].double() * cs[None, :].double())
    for Cb, Ct in res:
        if outs != "transposed":
            assert float((Cb[:, :N].double() - ref).abs().max() / ref.abs().max()) < 6e-3       # bf16 rounding
            assert bool((Cb[:, N:] == 7.0).all())                                              # padding untouched
        if outs != "rowmajor":
            assert float((Ct[:, :M].double() - ref.T).abs().max() / ref.abs().max()) < 6e-3
            assert bool((Ct[:, M:] == 7.0).all())
    if have_aids:
        assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])


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
