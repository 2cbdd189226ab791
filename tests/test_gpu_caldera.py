"""GPU parity of caldera() (cb_caldera_layer) against golden runs of the reference, the
numpy oracle, and size-independent properties at the BASELINE shape.

Tolerances (see DESIGN.md "Parity"): integer codes and scales are bit-exact wherever the
quantiser input is bit-identical (iterate 0 with the reference's global_scale injected);
the Hessian-weighted relative error of the returned iterate is within 1e-3 relative of the
reference (north_star); the rest of the trajectory is path dependent (the 2-bit Q update is
discontinuous) and is compared with a looser bound."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from ee274_convexcaldera_llm_quantization_b200 import _lib
from oracle import caldera_oracle as orc
from src.caldera.utils.dataclasses import CalderaParams
from src.caldera.utils.quantization import QuantizerFactory, unpack_codes
from src.caldera.decomposition.alg import caldera

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _tdiv(x, lv):
    """True IEEE division on the GPU (torch turns tensor / python_scalar into a multiply by the
    reciprocal on CUDA, which is not what the CPU reference computes)."""
    return x / torch.full((), float(lv), device=x.device)

FILES = sorted(glob.glob(os.path.join(GOLDEN, "caldera_*.npz")))
IDS = [os.path.basename(p)[8:-4] for p in FILES]


def _load(path):
    z = np.load(path)
    kw = json.loads(str(z["params"]))
    H = None
    if "H" in z:
        H = torch.from_numpy(z["H"])
    elif "h" in z:
        H = torch.diag(torch.from_numpy(z["h"]))      # dense diag_embed, as main.py:165 passes it
    return z, kw, H


def _params(kw):
    return CalderaParams(quant_factory_Q=QuantizerFactory(method="uniform", block_size=64),
                         quant_factory_LR=QuantizerFactory(method="uniform", block_size=64), **kw)


def _weighted_error_torch(W, h, Q, L, R):
    E = (Q + L @ R - W).double()
    hd = torch.ones(W.shape[1], dtype=torch.float64, device=W.device) if h is None else h.double()
    return float(((E * E * hd[None, :]).sum() / (W.double() ** 2 * hd[None, :]).sum()).sqrt())


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_golden(path):
    z, kw, H = _load(path)
    p = _params(kw)
    scale_W = bool(z["scale_W"])
    gs = float(z["global_scale"]) if scale_W else None
    d = caldera(p, torch.from_numpy(z["W"]), H, device=DEV, use_tqdm=False, scale_W=scale_W, global_scale=gs)
    m, n = z["W"].shape
    # ---- structure (alg.py:71-81, 108-111)
    assert d.Q.shape == (m, n) and d.L.shape == (m, p.rank) and d.R.shape == (p.rank, n)
    assert d.Q.dtype == d.L.dtype == d.R.dtype == torch.float32 and d.Q.is_cuda
    assert d.W.device.type == "cpu" and d.W.shape == (m, n)
    assert set(d.errors) == set(p.update_order)
    assert isinstance(d.global_scale, float) or d.global_scale == 1
    if scale_W:
        assert abs(d.global_scale - float(z["global_scale"])) <= 1e-7 * float(z["global_scale"])
    quantised = p.compute_low_rank_factors and (p.L_bits < 16 or p.R_bits < 16)
    # ---- trajectories
    ref_all, got_all = [], []
    for k in p.update_order:
        ref, got = z[f"errors_{k}"], np.array(d.errors[k])
        assert got.shape == ref.shape and np.isfinite(got).all()
        ref_all.append(ref)
        got_all.append(got)
        # (measured, scripts/golden_deviation.py: <= 4.3e-4 with fp32 factors; <= 4.5e-2 with re-quantised factors on
        #  these 96 x 128 ... 512 x 384 matrices, where one flipped 4-bit code moves the trajectory; 1.3e-2 for rand_svd)
        loose = 6e-2 if (quantised or p.rand_svd) else 5e-3
        np.testing.assert_allclose(got, ref, rtol=loose)
    ref_all, got_all = np.concatenate(ref_all), np.concatenate(got_all)
    # first sub-step: quantiser input (or SVD input) is bit-identical to the reference's
    first = p.update_order[0]
    np.testing.assert_allclose(d.errors[first][0], z[f"errors_{first}"][0], rtol=2e-5 if first == "Q" else 1e-3)
    # ---- the returned (best) iterate: north_star tolerance 1e-3 relative
    order = p.update_order
    k = len(order)
    seq_ref = [float(z[f"errors_{order[s % k]}"][s // k]) for s in range(p.iters * k)]
    seq_got = [d.errors[order[s % k]][s // k] for s in range(p.iters * k)]
    ref_best, got_best = min(seq_ref[k - 1:]), min(seq_got[k - 1:])   # strict arg-min once all updated (alg.py:105)
    # north_star: within 1e-3 relative of the reference (one-sided: a lower error is never a failure;
    # the lower guard only catches a broken metric).  With re-quantised factors or rand_svd the
    # reference itself is only reproducible to a few percent.
    # (measured: <= 1.5e-5 with fp32 factors, <= 4.5e-3 with quantised factors; the LPLR iteration itself is pinned
    #  stage by stage in tests/test_gpu_lplr_stage.py)
    tol = 1e-3 if not (quantised or p.rand_svd) else 1e-2
    assert got_best <= ref_best * (1 + tol), (got_best, ref_best)
    assert got_best >= ref_best * (1 - max(10 * tol, 2e-2)), (got_best, ref_best)
    # ---- self consistency: the error reported for the returned iterate is the error of the returned tensors
    if "H" in z:
        Hd = H.to(DEV).double()
        Hd = (Hd + Hd.T) / 2
        if kw.get("sigma_reg", 0) and p.activation_aware_LR:
            lam_min = float(torch.linalg.eigvalsh(Hd)[0])
            if lam_min < kw["sigma_reg"]:                       # alg.py:59-63
                Hd = Hd + (kw["sigma_reg"] - lam_min) * torch.eye(Hd.shape[0], device=DEV, dtype=torch.float64)
        Wd64 = d.W.to(DEV).double()
        E64 = (d.Q + d.L @ d.R).double() - Wd64
        consistent = float((torch.trace(E64 @ Hd @ E64.T) / torch.trace(Wd64 @ Hd @ Wd64.T)).sqrt())
    else:
        h = None if H is None else torch.diagonal(H).to(DEV)
        if kw.get("sigma_reg", 0) and p.activation_aware_LR and h is not None and float(h.min()) < kw["sigma_reg"]:
            h = h + (kw["sigma_reg"] - float(h.min()))
        consistent = _weighted_error_torch(d.W.to(DEV), h, d.Q, d.L, d.R)
    assert d.best_step == int(np.argmin(seq_got[k - 1:])) + k - 1
    np.testing.assert_allclose(consistent, seq_got[d.best_step], rtol=5e-5)
    # ---- codes
    if p.compute_quantized_component:
        assert d.Q_idxs.shape == (1, m * n) and d.Q_idxs.dtype == torch.int8
        assert d.Q_scale.shape == (1, 1)
        lv = 2 ** (p.Q_bits - 1) - 1
        assert torch.equal(d.Q, (_tdiv(d.Q_idxs.float(), lv) * d.Q_scale).reshape(m, n))
        assert torch.equal(unpack_codes(d.Q_packed, p.Q_bits, m * n), d.Q_idxs.reshape(-1))
        if "Q_idxs" in z and not quantised and not p.rand_svd:
            match = float((d.Q_idxs.cpu().numpy() == z["Q_idxs"]).mean())
            assert match > 0.99, match
    if quantised:
        assert d.L_idxs.shape == (1, p.rank * m) and d.R_idxs.shape == (1, p.rank * n)
        lvl, lvr = 2 ** (p.L_bits - 1) - 1, 2 ** (p.R_bits - 1) - 1
        Lq = (_tdiv(d.L_idxs.float(), lvl) * d.L_scale).reshape(p.rank, m).T     # codes of L.T (alg.py:171)
        Rq = (_tdiv(d.R_idxs.float(), lvr) * d.R_scale).reshape(p.rank, n)
        assert torch.equal(Lq.contiguous(), d.L) and torch.equal(Rq, d.R)
        assert int(d.L_idxs.abs().max()) == lvl and int(d.R_idxs.abs().max()) == lvr


def test_iterate0_codes_bit_exact():
    """Q-first, one sub-step: residual == W exactly, so codes and scale must equal the reference's."""
    z, kw, H = _load(os.path.join(GOLDEN, "caldera_q_only.npz"))
    p = _params(kw)
    d = caldera(p, torch.from_numpy(z["W"]), H, device=DEV, use_tqdm=False, global_scale=float(z["global_scale"]))
    np.testing.assert_array_equal(d.Q_idxs.cpu().numpy(), z["Q_idxs"])
    np.testing.assert_array_equal(d.Q_scale.cpu().numpy(), z["Q_scale"])
    np.testing.assert_array_equal(d.Q.cpu().numpy(), z["Q"])
    np.testing.assert_array_equal(d.W.numpy(), z["W_scaled"])
    np.testing.assert_allclose(d.errors["Q"], z["errors_Q"], rtol=2e-6)
    # independently computed global_scale: within an ulp or two of torch's reduction
    d2 = caldera(p, torch.from_numpy(z["W"]), H, device=DEV, use_tqdm=False)
    assert abs(d2.global_scale - float(z["global_scale"])) <= 2.5e-7 * float(z["global_scale"])


def test_h_forms_agree():
    """None / 1-D diagonal / dense diag_embed Hessians take the same device path."""
    g = torch.Generator().manual_seed(5)
    W = 0.02 * torch.randn(128, 192, generator=g)
    h = 0.5 + torch.rand(192, generator=g)
    p = _params(dict(Q_bits=2, L_bits=16, R_bits=16, rank=8, iters=2, update_order=["Q", "LR"]))
    a = caldera(p, W, h, device=DEV, use_tqdm=False, seed=3)
    b = caldera(p, W.to(DEV), torch.diag(h).to(DEV), device=DEV, use_tqdm=False, seed=3)
    assert a.errors == b.errors and torch.equal(a.Q_idxs, b.Q_idxs) and torch.equal(a.L, b.L)
    c = caldera(p, W, None, device=DEV, use_tqdm=False, seed=3)
    e = caldera(p, W, torch.ones(192), device=DEV, use_tqdm=False, seed=3)
    assert c.errors == e.errors


def test_error_conventions_gpu():
    W = torch.randn(64, 64)
    with pytest.raises(AttributeError):
        caldera(_params(dict(L_bits=4, R_bits=4, rank=4, iters=1, lplr_iters=0, update_order=["Q", "LR"])),
                W, device=DEV, use_tqdm=False)
    with pytest.raises(AssertionError):
        caldera(_params(dict(Q_bits=3, rank=4, iters=1, update_order=["Q"])), W, device=DEV, use_tqdm=False)
    # update_order == [] (the dataclass default): nothing runs, zeros come back (dataclasses.py:48-57)
    d = caldera(_params(dict(rank=4, iters=2)), W, device=DEV, use_tqdm=False)
    assert d.errors == {} and float(d.Q.abs().max()) == 0 and float(d.L.abs().max()) == 0
    assert d.Q_idxs is None and d.Q_scale == 1


@pytest.mark.parametrize("lbits", [16, 4])
def test_vs_oracle_medium(lbits):
    g = torch.Generator().manual_seed(77)
    m, n, r = 320, 448, 24
    W = 0.02 * torch.randn(m, n, generator=g)
    h = torch.exp(0.8 * torch.randn(n, generator=g))
    kw = dict(Q_bits=4, L_bits=lbits, R_bits=lbits, rank=r, iters=3, lplr_iters=3, update_order=["LR", "Q"])
    ref = orc.caldera_oracle(orc.OracleParams(**kw), W.numpy(), h.numpy())
    d = caldera(_params(kw), W, h, device=DEV, use_tqdm=False, global_scale=ref.global_scale)
    np.testing.assert_allclose(d.errors["LR"][0], ref.errors["LR"][0], rtol=1e-3 if lbits == 16 else 2e-2)
    best_ref = min(ref.errors["Q"])
    best_got = min(d.errors["Q"])
    # lbits == 4: whole-tensor 4-bit re-quantisation of L/R makes the trajectory chaotic; the numpy
    # device model spreads over 0.170..0.190 across sketch seeds for this very input
    # lbits == 16: 4-bit Q with a heavy-tailed h is itself a chaotic iteration -- the numpy device model
    # spreads over 0.1434..0.1447 (+-5e-3) across sketch seeds whatever the power-iteration count
    tol = 8e-3 if lbits == 16 else 8e-2
    assert best_got <= best_ref * (1 + tol) and best_got >= best_ref * (1 - 10 * tol), (best_got, best_ref)


@pytest.mark.parametrize("cfg", ["c2_lr16", "c4_lr4"])
def test_full_size_properties(cfg):
    """BASELINE config 2 shape (4096 x 4096, rank 128, Q 2-bit): invariants that need no CPU oracle."""
    m = n = 4096
    r = 128
    g = torch.Generator().manual_seed(1000)
    W = 0.02 * torch.randn(m, n, generator=g)
    h = 0.5 + torch.rand(n, generator=g)
    lb = 16 if cfg == "c2_lr16" else 4
    kw = dict(Q_bits=2, L_bits=lb, R_bits=lb, rank=r, iters=2 if lb == 4 else 3, lplr_iters=2,
              update_order=["Q", "LR"])
    d = caldera(_params(kw), W, h, device=DEV, use_tqdm=False, W_copy="device")
    eq, elr = d.errors["Q"], d.errors["LR"]
    assert 0.97 < eq[0] < 0.995            # survey probe: first-Q error 0.987 at 4096^2
    assert elr[0] < eq[0] and min(elr) < 0.95
    assert d.best_step >= 1
    hd = h.to(DEV)
    consistent = _weighted_error_torch(d.W, hd, d.Q, d.L, d.R)
    seq = [x for pair in zip(eq, elr) for x in pair]
    np.testing.assert_allclose(consistent, seq[d.best_step], rtol=2e-5)
    assert abs(consistent - min(seq[1:])) < 1e-6
    assert torch.equal(d.Q, (_tdiv(d.Q_idxs.float(), 1) * d.Q_scale).reshape(m, n))
    assert torch.equal(unpack_codes(d.Q_packed, 2, m * n), d.Q_idxs.reshape(-1))
    # rank-r optimality of the first LR step against the exact spectrum (torch SVD on the GPU)
    if lb == 16:
        Wd = d.W
        amax = Wd.abs().max()
        Q0 = torch.round(Wd / amax) * amax
        S = torch.linalg.svdvals(((Wd - Q0) * hd.sqrt()[None, :]).double())
        opt = float(((S[r:] ** 2).sum() / ((Wd.double() ** 2) * hd[None, :].double()).sum()).sqrt())
        assert elr[0] <= opt * (1 + 1e-3), (elr[0], opt)
        # bf16 operands: the basis is orthonormal to bf16 rounding (2^-8), not to fp32
        assert float((d.L.T @ d.L - torch.eye(r, device=DEV)).abs().max()) < 1e-2 or d.best_step != 1


def test_tensor_core_path_agrees_with_simt():
    """bf16 tcgen05 contractions vs fp32 SIMT contractions on the same layer (aligned shape)."""
    g = torch.Generator().manual_seed(11)
    m, n, r = 1024, 768, 32
    W = 0.02 * torch.randn(m, n, generator=g)
    h = torch.exp(0.5 * torch.randn(n, generator=g))
    kw = dict(Q_bits=4, L_bits=16, R_bits=16, rank=r, iters=3, update_order=["Q", "LR"])
    a = caldera(_params(kw), W, h, device=DEV, use_tqdm=False, seed=5, use_tensor_cores=True)
    b = caldera(_params(kw), W, h, device=DEV, use_tqdm=False, seed=5, use_tensor_cores=False)
    assert a.device_stats["tc_watchdog"] == 0
    assert a.errors["Q"][0] == b.errors["Q"][0]                     # no contraction involved yet
    np.testing.assert_allclose(a.errors["LR"][0], b.errors["LR"][0], rtol=1e-3)   # same Q, same residual
    np.testing.assert_allclose(a.errors["LR"], b.errors["LR"], rtol=5e-3)         # then path dependent
    assert float((b.L.T @ b.L - torch.eye(r, device=DEV)).abs().max()) < 1e-3 or b.best_step % 2 == 0
    hd = h.to(DEV)
    np.testing.assert_allclose(_weighted_error_torch(a.W.to(DEV), hd, a.Q, a.L, a.R),
                               [e for pair in zip(a.errors["Q"], a.errors["LR"]) for e in pair][a.best_step],
                               rtol=5e-5)


@pytest.mark.parametrize("n,kind", [(48, "lowrank"), (160, "lowrank"), (512, "lowrank"), (96, "full"), (640, "indefinite"),
                                    (2048, "lowrank")])
def test_dense_hessian_min_eig_shift(n, kind):
    """alg.py:57-64 for a dense H: lambda_min by Lanczos on the device against torch.linalg.eigvalsh, and the shift
    H += (sigma_reg - lambda_min) I.  (The golden cases *_dense_sigma_reg run the whole decomposition with it.)"""
    lib = _lib.load()
    g = torch.Generator().manual_seed(n)
    if kind == "lowrank":
        X = torch.randn(n // 2, n, generator=g) * (0.25 + torch.rand(n, generator=g))[None, :]
        H = X.T @ X / (n // 2)
    elif kind == "full":
        X = torch.randn(4 * n, n, generator=g)
        H = X.T @ X / (4 * n)
    else:
        A = torch.randn(n, n, generator=g)
        H = (A + A.T) / (2 * n ** 0.5)                      # Wigner matrix: lambda_min ~ -sqrt(2)
    lam = torch.linalg.eigvalsh(H.double())
    lam_min, lam_max = float(lam[0]), float(lam[-1])
    for sigma_reg in (0.0, 1e-3, 0.5):
        Hd = H.to(DEV).contiguous()
        ws = torch.empty(lib.cb_min_eig_shift_workspace_bytes(n), dtype=torch.uint8, device=DEV)
        stats = torch.zeros(2, device=DEV)
        _lib.check(lib.cb_min_eig_shift_f32(_lib.ptr(Hd), n, sigma_reg, _lib.ptr(stats), _lib.ptr(ws), ws.numel(),
                                            _lib.stream_ptr()), "min_eig_shift")
        shift, est = (float(x) for x in stats.tolist())
        # a separated lower end (rank-deficient Gram matrix, small n) is found to fp32 accuracy; the edge of a
        # continuous spectrum converges like range / k^2 from above
        tol = 2e-5 * lam_max if kind == "lowrank" or n <= 96 else 2e-2 * (lam_max - lam_min)
        assert abs(est - lam_min) <= tol + 1e-6, (kind, n, est, lam_min)
        assert est >= lam_min - 1e-5 * max(abs(lam_max), 1.0)
        want = max(0.0, sigma_reg - est)
        np.testing.assert_allclose(shift, want, rtol=1e-6, atol=1e-9)
        assert torch.allclose(Hd.cpu(), H + shift * torch.eye(n), atol=1e-6 * max(1.0, lam_max))


def test_cuda_graph_replay_matches_eager():
    """caldera(use_cuda_graph=True) replays a captured graph of the layer: same numbers as the eager
    launch sequence, for different inputs and seeds through the same cached graph.  (fp32 SIMT contractions: with
    tensor cores the graph path runs the batched driver, whose contraction kernel sums in a different order --
    tests/test_gpu_batch.py compares those two.)"""
    g = torch.Generator().manual_seed(21)
    kw = dict(Q_bits=2, L_bits=4, R_bits=4, rank=16, iters=2, lplr_iters=2, update_order=["Q", "LR"])
    for trial in range(3):
        W = 0.02 * torch.randn(512, 384, generator=g)
        h = 0.5 + torch.rand(384, generator=g)
        a = caldera(_params(kw), W, h, device=DEV, use_tqdm=False, seed=trial, use_cuda_graph=False, use_tensor_cores=False)
        b = caldera(_params(kw), W.pin_memory(), h, device=DEV, use_tqdm=False, seed=trial, use_cuda_graph=True,
                    use_tensor_cores=False)
        assert a.errors == b.errors
        assert torch.equal(a.Q_idxs, b.Q_idxs) and torch.equal(a.L, b.L) and torch.equal(a.R_idxs, b.R_idxs)
        assert torch.equal(a.Q_packed, b.Q_packed) and torch.equal(a.W, b.W)
        assert a.best_step == b.best_step and a.global_scale == b.global_scale


@pytest.mark.parametrize("tc", [True, False], ids=["tcgen05", "simt"])
@pytest.mark.parametrize("lbits", [16, 4])
def test_repeated_runs_are_bitwise_identical(tc, lbits):
    """a layer's result depends only on (W, H, params, seed) -- not on which GPU, stream or
    neighbour it ran beside.  Split-K contractions sum their slices in a fixed order (no floating-point
    atomics), so repeated runs agree bit for bit even while another stream perturbs CTA scheduling."""
    g = torch.Generator().manual_seed(31)
    m, n, r = 2048, 1536, 64
    W = 0.02 * torch.randn(m, n, generator=g)
    h = 0.5 + torch.rand(n, generator=g)
    kw = dict(Q_bits=2, L_bits=lbits, R_bits=lbits, rank=r, iters=2, lplr_iters=2, update_order=["Q", "LR"])
    first = caldera(_params(kw), W, h, device=DEV, use_tqdm=False, seed=9, use_tensor_cores=tc)
    noise_stream = torch.cuda.Stream()
    X = torch.randn(4096, 4096, device=DEV)
    for trial in range(3):
        with torch.cuda.stream(noise_stream):
            for _ in range(8 * (trial + 1)):
                X = torch.tanh(X @ X) * 0.01
        again = caldera(_params(kw), W, h, device=DEV, use_tqdm=False, seed=9, use_tensor_cores=tc)
        assert again.errors == first.errors
        assert torch.equal(again.Q_idxs, first.Q_idxs) and torch.equal(again.L, first.L) and torch.equal(again.R, first.R)
        assert again.best_step == first.best_step
    torch.cuda.synchronize()


def _fullsize_layer(m, n, seed):
    g = torch.Generator().manual_seed(seed)
    W = 0.02 * torch.randn(m, n, generator=g, dtype=torch.float32)
    h = 0.5 + torch.rand(n, generator=g, dtype=torch.float32)
    return W, h


def test_fullsize_golden_headline_config():
    """BASELINE config 2 at full size (4096 x 4096, rank 128, Q 2-bit, 5 iterations) against the UNMODIFIED
    reference run on CPU (tests/golden/make_golden_fullsize.py): global_scale, iterate-0 codes bit for bit (SHA-256
    of Q_idxs) and scale, the best error within the 1e-3 relative bar of the north star, the Q scale of the best
    iterate, and the exact-match fraction of the best iterate's packed codes."""
    from ee274_convexcaldera_llm_quantization_b200 import parity
    z, _ = parity.load_fullsize_golden("c2")
    W, h = _fullsize_layer(4096, 4096, 1000)
    kw = dict(Q_bits=2, L_bits=16, R_bits=16, rank=128, iters=5, update_order=["Q", "LR"])
    # iterate 0: the quantiser input is W / global_scale itself -> bit-exact codes and scale
    d0 = caldera(_params(dict(kw, iters=1, update_order=["Q"])), W, h, device=DEV, use_tqdm=False, W_copy="none",
                 global_scale=z["global_scale"])
    assert parity.codes_sha256(d0.Q_idxs.reshape(-1)) == z["q_idxs_iter0_sha256"]
    assert np.float32(float(d0.Q_scale)) == np.float32(z["Q_scale_iter0"])
    for mode in ("latency", "throughput"):
        try:
            _lib.set_execution_mode(mode)
            d = caldera(_params(kw), W, h, device=DEV, use_tqdm=False, W_copy="none", global_scale=z["global_scale"])
            d_own = caldera(_params(kw), W, h, device=DEV, use_tqdm=False, W_copy="none")
        finally:
            _lib.set_execution_mode("latency")
        np.testing.assert_allclose(d_own.global_scale, z["global_scale"], rtol=2e-7)
        np.testing.assert_allclose(d.errors["Q"][0], z["errors"]["Q"][0], rtol=2e-6)     # no rank-r step involved yet
        np.testing.assert_allclose(d.errors["LR"][0], z["errors"]["LR"][0], rtol=1e-3)   # first rank-r step vs exact SVD
        flat = [e for pair in zip(d.errors["Q"], d.errors["LR"]) for e in pair]
        best = min(flat[1:])
        assert abs(best - z["best_error"]) <= 1e-3 * z["best_error"], (mode, best, z["best_error"])
        assert flat[d.best_step] == best
        rep = parity.parity_report(d, "c2", d0)
        print(f"\n[parity c2, {mode}] {rep}")
        assert rep["code_match_iter0"] == 1.0 and rep["q_scale_iter0_equal"]
        # the reference's best iterate is step 1 (the first LR update), whose Q is still the iterate-0 Q: if ours
        # is too, its codes and scale are bit-identical; otherwise the mismatch rate is what is reported
        if d.best_step == z["best_step"] == 1:
            assert rep["code_match_best"] == 1.0 and rep["q_scale_rel_diff"] == 0.0
        else:
            assert rep["code_match_best"] >= 0.99 and rep["q_scale_rel_diff"] <= 1e-2


@pytest.mark.parametrize("name,m,n,seed", [("c3", 11008, 4096, 1004), ("c3t", 4096, 11008, 1005)])
def test_fullsize_golden_config3_quantised_factors(name, m, n, seed):
    """SURVEY configuration 3 at full size, both orientations (11008 x 4096 and 4096 x 11008, rank 256, Q 2-bit,
    L/R 4-bit, 5 outer and 5 LPLR iterations) against the UNMODIFIED reference on CPU."""
    from ee274_convexcaldera_llm_quantization_b200 import parity
    z, _ = parity.load_fullsize_golden(name)
    W, h = _fullsize_layer(m, n, seed)
    kw = dict(Q_bits=2, L_bits=4, R_bits=4, rank=256, iters=5, lplr_iters=5, update_order=["Q", "LR"])
    d0 = caldera(_params(dict(kw, iters=1, update_order=["Q"])), W, h, device=DEV, use_tqdm=False, W_copy="none",
                 global_scale=z["global_scale"])
    assert parity.codes_sha256(d0.Q_idxs.reshape(-1)) == z["q_idxs_iter0_sha256"]
    assert np.float32(float(d0.Q_scale)) == np.float32(z["Q_scale_iter0"])
    d_own = caldera(_params(dict(kw, iters=0)), W, h, device=DEV, use_tqdm=False, W_copy="none")
    np.testing.assert_allclose(d_own.global_scale, z["global_scale"], rtol=2e-7)
    d = caldera(_params(kw), W, h, device=DEV, use_tqdm=False, W_copy="none", global_scale=z["global_scale"])
    np.testing.assert_allclose(d.errors["Q"][0], z["errors"]["Q"][0], rtol=2e-6)
    flat = [e for pair in zip(d.errors["Q"], d.errors["LR"]) for e in pair]
    best = min(flat[1:])
    rep = parity.parity_report(d, name, d0)
    print(f"\n[parity {name}] {rep}; trajectory {flat}; reference {z['errors']}")
    # 4-bit whole-tensor re-quantisation of L and R: bf16 contractions in the least-squares updates move single
    # codes, and the reference's own trajectory is reproducible to ~3e-3 only
    assert rep["best_err_rel_diff"] <= 2e-3, rep
    assert rep["code_match_iter0"] == 1.0 and rep["code_match_best"] >= 0.99
    # the factor scales are abs-max values (one heavy-tailed element each) of whichever inner iterate won
    np.testing.assert_allclose(float(d.L_scale), z["L_scale"], rtol=0.15)
    np.testing.assert_allclose(float(d.R_scale), z["R_scale"], rtol=0.35)
    assert d.L_idxs.shape == (1, m * 256) and d.R_idxs.shape == (1, 256 * n)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_second_device_after_first_in_one_process():
    """kernel attribute opt-ins (> 48 KiB dynamic shared memory) are per device; a process that has already
    used cuda:0 must be able to decompose on cuda:1 (single-layer and batched drivers, packed consumer)."""
    from ee274_convexcaldera_llm_quantization_b200.alg import caldera_async
    g = torch.Generator().manual_seed(77)
    W = 0.02 * torch.randn(512, 768, generator=g)
    h = 0.5 + torch.rand(768, generator=g)
    fac = QuantizerFactory(method="uniform", block_size=64)
    qp = CalderaParams(Q_bits=2, L_bits=4, R_bits=4, rank=32, iters=2, lplr_iters=2, update_order=["Q", "LR"],
                       quant_factory_Q=fac, quant_factory_LR=fac)
    out = {}
    for dev in ("cuda:0", "cuda:1"):
        d = caldera(qp, W, h, device=dev, use_tqdm=False, seed=3)
        hs = [caldera_async(qp, W, h, device=dev, use_tqdm=False, seed=3, batch_hint=2) for _ in range(2)]
        b = [x.result() for x in hs]
        assert str(d.Q.device) == dev and str(b[0].L.device) == dev
        out[dev] = (d.errors, b[0].errors, d.Q_idxs.cpu(), b[1].Q_idxs.cpu())
    assert out["cuda:0"][0] == out["cuda:1"][0] and out["cuda:0"][1] == out["cuda:1"][1]
    assert torch.equal(out["cuda:0"][2], out["cuda:1"][2]) and torch.equal(out["cuda:0"][3], out["cuda:1"][3])
