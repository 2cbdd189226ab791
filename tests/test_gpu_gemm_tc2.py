"""The batched, persistent CTA-pair tcgen05 GEMM (csrc/gemm_tc2.cu: cta_group::2, two TMEM accumulators) against
torch on bf16-rounded operands: every output form, ragged shapes, several tiles per cluster (both accumulators and
several laps of the shared-memory ring), batch strides."""
import pytest
import torch

from ee274_convexcaldera_llm_quantization_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _run(batch, M, N, K, max_clusters=0, outs=("C", "Cb", "Ct"), scales=False, alpha=1.0, pad=0, dynamic=True):
    lib = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(batch + 3 * M + 5 * N + 7 * K + max_clusters)
    lda = (K + 7) // 8 * 8 + pad
    A = torch.full((batch, M, lda), 7.0, device=DEV, dtype=torch.bfloat16)       # padding beyond K must never be read
    B = torch.full((batch, N, lda), -5.0, device=DEV, dtype=torch.bfloat16)
    A[:, :, :K] = torch.randn(batch, M, K, generator=g, device=DEV).bfloat16()
    B[:, :, :K] = torch.randn(batch, N, K, generator=g, device=DEV).bfloat16()
    C = torch.full((batch, M, N), float("nan"), device=DEV) if "C" in outs else None
    Cb = torch.full((batch, M, N), float("nan"), device=DEV, dtype=torch.bfloat16) if "Cb" in outs else None
    Ct = torch.full((batch, N, M), float("nan"), device=DEV, dtype=torch.bfloat16) if "Ct" in outs else None
    cs = (0.5 + torch.rand(batch, N, generator=g, device=DEV)) if scales else None
    rs = (0.5 + torch.rand(batch, M, generator=g, device=DEV)) if scales else None
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    counter = torch.zeros(2, dtype=torch.int32, device=DEV) if dynamic else None
    sb = lambda t: 0 if t is None else t.stride(0) * t.element_size()  # noqa: E731
    st = lib.cb_gemm_bf16_tn_batched(batch, M, N, K, alpha, _lib.ptr(A), lda, sb(A), _lib.ptr(B), lda, sb(B),
                                     _lib.ptr(C), N, sb(C), _lib.ptr(Cb), N, sb(Cb), _lib.ptr(Ct), M, sb(Ct),
                                     _lib.ptr(cs), sb(cs), _lib.ptr(rs), sb(rs), max_clusters, _lib.ptr(counter), _lib.ptr(flag),
                                     _lib.stream_ptr())
    _lib.check(st, "gemm_bf16_tn_batched")
    torch.cuda.synchronize()
    assert int(flag.item()) == 0, "pipeline watchdog fired"
    if counter is not None:
        assert counter.tolist() == [0, 0], "the dynamic tile scheduler must leave its counters zeroed"
    ref = alpha * torch.bmm(A[:, :, :K].double(), B[:, :, :K].double().transpose(1, 2))
    if scales:
        ref = ref * rs.double()[:, :, None] * cs.double()[:, None, :]
    scale = ref.abs().max().clamp_min(1e-30)
    errs = {}
    if C is not None:
        errs["C"] = float((C.double() - ref).abs().max() / scale)
    if Cb is not None:
        errs["Cb"] = float((Cb.double() - ref).abs().max() / scale)
    if Ct is not None:
        errs["Ct"] = float((Ct.double().transpose(1, 2) - ref).abs().max() / scale)
    return errs


def _check(errs):
    for k, e in errs.items():
        assert e < (2e-5 if k == "C" else 5e-3), (k, e)        # bf16 outputs carry 8 mantissa bits


@pytest.mark.parametrize("batch,M,N,K", [
    (1, 256, 224, 512), (1, 256, 256, 64), (1, 512, 224, 4096), (3, 1024, 224, 1024), (2, 300, 100, 200),
    (1, 128, 32, 128), (2, 256, 512, 256), (1, 1000, 333, 776), (4, 224, 224, 2048), (1, 4096, 224, 4096), (1, 37, 53, 32)])
def test_gemm_tc2_matches_torch(batch, M, N, K):
    _check(_run(batch, M, N, K))


@pytest.mark.parametrize("max_clusters", [1, 2, 3, 5])
def test_gemm_tc2_many_tiles_per_cluster(max_clusters):
    """Few clusters, many tiles: both TMEM accumulators, their full/empty barriers and many laps of the smem ring."""
    _check(_run(3, 1280, 224, 640, max_clusters=max_clusters))
    _check(_run(2, 768, 600, 192, max_clusters=max_clusters))
    _check(_run(3, 1280, 224, 640, max_clusters=max_clusters, dynamic=False))     # fixed tile -> cluster assignment


def test_gemm_tc2_scales_alpha_and_output_subsets():
    _check(_run(2, 512, 224, 512, scales=True, alpha=0.25))
    _check(_run(2, 512, 224, 512, outs=("Cb",)))
    _check(_run(2, 512, 224, 512, outs=("Ct",), scales=True))
    _check(_run(1, 512, 224, 512, outs=("C",), pad=16))


def test_gemm_tc2_is_bitwise_reproducible():
    lib = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(9)
    A = torch.randn(4, 1024, 512, generator=g, device=DEV).bfloat16()
    B = torch.randn(4, 224, 512, generator=g, device=DEV).bfloat16()
    outs = []
    counter = torch.zeros(2, dtype=torch.int32, device=DEV)
    for mc, dyn in ((0, True), (3, True), (0, False), (5, True)):
        C = torch.empty(4, 1024, 224, device=DEV)
        flag = torch.zeros(1, dtype=torch.int32, device=DEV)
        _lib.check(lib.cb_gemm_bf16_tn_batched(4, 1024, 224, 512, 1.0, _lib.ptr(A), 512, A.stride(0) * 2, _lib.ptr(B), 512,
                                               B.stride(0) * 2, _lib.ptr(C), 224, C.stride(0) * 4, None, 0, 0, None, 0, 0,
                                               None, 0, None, 0, mc, _lib.ptr(counter) if dyn else None, _lib.ptr(flag),
                                               _lib.stream_ptr()), "g2")
        torch.cuda.synchronize()
        assert int(flag.item()) == 0
        outs.append(C)
    assert all(torch.equal(outs[0], o) for o in outs[1:])    # no dependence on the tile -> cluster map


@pytest.mark.parametrize("batch,M,N,K,pad", [(1, 512, 512, 384, 0), (3, 1024, 1024, 96, 0), (2, 300, 100, 200, 0), (1, 37, 52, 32, 0),
                                             (2, 1000, 332, 776, 0), (1, 256, 224, 512, 8), (4, 4096, 256, 384, 0)])
def test_gemm_tc2_fp32_only_output_goes_through_tma_stores(batch, M, N, K, pad):
    """fp32 C as the only output (the L R product of the batched driver): staged per warp in swizzled shared memory and
    written with cp.async.bulk.tensor stores; ragged edges are clipped by the tensor map, padding columns of C (ldc >
    N) and the rows of the next batch item must stay untouched."""
    lib = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    A = torch.randn(batch, M, (K + 7) // 8 * 8, generator=g, device=DEV).bfloat16()
    B = torch.randn(batch, N, (K + 7) // 8 * 8, generator=g, device=DEV).bfloat16()
    A[:, :, K:] = 0
    B[:, :, K:] = 0
    ldc = N + pad
    C = torch.full((batch, M + 3, ldc), -777.0, device=DEV)                      # 3 guard rows per item, `pad` guard columns
    cs = 0.5 + torch.rand(batch, N, generator=g, device=DEV)
    rs = 0.5 + torch.rand(batch, M, generator=g, device=DEV)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    counter = torch.zeros(2, dtype=torch.int32, device=DEV)
    _lib.check(lib.cb_gemm_bf16_tn_batched(batch, M, N, K, 0.5, _lib.ptr(A), A.shape[2], A.stride(0) * 2, _lib.ptr(B), B.shape[2],
                                           B.stride(0) * 2, _lib.ptr(C), ldc, C.stride(0) * 4, None, 0, 0, None, 0, 0,
                                           _lib.ptr(cs), cs.stride(0) * 4, _lib.ptr(rs), rs.stride(0) * 4, 0, _lib.ptr(counter),
                                           _lib.ptr(flag), _lib.stream_ptr()), "g2")
    torch.cuda.synchronize()
    assert int(flag.item()) == 0
    ref = 0.5 * torch.bmm(A.double(), B.double().transpose(1, 2)) * rs.double()[:, :, None] * cs.double()[:, None, :]
    err = float((C[:, :M, :N].double() - ref).abs().max() / ref.abs().max())
    assert err < 2e-5, err
    assert bool((C[:, M:, :] == -777.0).all()) and bool((C[:, :, N:] == -777.0).all())
