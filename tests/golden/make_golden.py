"""Generate golden vectors by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Writes tests/golden/quantizer.npz and tests/golden/caldera_<case>.npz.  The fixtures pin
`oracle/caldera_oracle.py` (tests/test_oracle_golden.py) and are compared directly with
the CUDA path in the `-m gpu` tests.  Valid for the torch build named in `meta`.
"""
import os
import sys
import json

import numpy as np
import torch

REF = "/root/reference/rank-constrained-regression-main"
sys.path.insert(0, REF)
from src.caldera.utils.quantization import QuantizerFactory, LowMemoryQuantizer  # noqa: E402
from src.caldera.utils.dataclasses import CalderaParams  # noqa: E402
from src.caldera.decomposition.alg import caldera  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def t2n(t):
    return t.detach().cpu().numpy()


# --------------------------------------------------------------------------- quantiser
def quantizer_cases():
    out = {}
    names = []
    xkeys = {}
    keep = []

    def run(name, x, bits, bs):
        q = QuantizerFactory(method="uniform", block_size=bs).get_quantizer(bits)
        codes, scales, shape = q.quantize_block(x)
        deq = q.dequantize_block(codes, scales, shape)
        keep.append(x)  # keep ids unique
        key = xkeys.setdefault(id(x), name)
        if key == name:
            out[f"{name}/x"] = t2n(x.contiguous())
        out[f"{name}/xref"] = np.array(key)
        out[f"{name}/codes"] = t2n(codes)
        out[f"{name}/scales"] = t2n(scales)
        out[f"{name}/deq"] = t2n(deq)
        out[f"{name}/bits_bs"] = np.array([bits, bs], dtype=np.int64)
        names.append(name)

    # known-answer rows from the survey probes
    run("kat2_a", torch.tensor([[2.0, 1.0, -1.0, 0.999]]), 2, 4)
    run("kat2_b", torch.tensor([[4.0, 2.0000002, -2.0, 1.9999999]]), 2, 4)
    run("kat2_zero", torch.zeros(1, 4), 2, 4)
    run("kat4_ties", torch.tensor([[7.0, 0.5, 1.5, 2.5]]), 4, 4)
    run("kat4_negties", torch.tensor([[-7.0, -0.5, -1.5, -2.5, 3.5, -3.5, 6.5, -6.5]]), 4, 8)
    run("kat8_ties", torch.tensor([[127.0, 0.5, 1.5, -2.5, 126.5, -126.5, 63.5, 64.5]]), 8, 8)

    g = torch.Generator().manual_seed(1234)
    x = torch.randn(128, 256, generator=g) * 0.02
    for bits in (2, 4, 8, 16):
        for bs in (64, 256, 128 * 256):
            run(f"rand_b{bits}_bs{bs}", x, bits, bs)
    # heavy-tailed values and exact zeros mixed in
    y = torch.randn(64, 192, generator=g)
    y = y * torch.exp(2.0 * torch.randn(64, 192, generator=g))
    y[3, :64] = 0.0
    y[10, 5] = 1e-12
    for bits in (2, 4, 8, 16):
        run(f"heavy_b{bits}_bs64", y, bits, 64)
    # non-contiguous input (the reference quantises L.T, alg.py:171)
    zt = torch.randn(48, 160, generator=g).T
    for bits in (2, 4):
        run(f"transposed_b{bits}_bs64", zt, bits, 64)
        run(f"transposed_b{bits}_whole", zt, bits, 48 * 160)
    # odd shape with block == whole tensor
    w = torch.randn(37, 53, generator=g)
    run("odd_b4_whole", w, 4, 37 * 53)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "quantizer.npz"), **out)
    print("quantizer:", len(names), "cases")


# --------------------------------------------------------------------------- caldera
def make_inputs(m, n, seed, hkind):
    g = torch.Generator().manual_seed(seed)
    W = 0.02 * torch.randn(m, n, generator=g)
    if hkind == "diag":
        H = torch.diag(0.5 + torch.rand(n, generator=g))
    elif hkind == "heavy":
        H = torch.diag(torch.exp(1.5 * torch.randn(n, generator=g)))
    elif hkind == "dense":
        X = torch.randn(4 * n, n, generator=g) * (0.25 + torch.rand(n, generator=g))[None, :]
        H = X.T @ X / (4 * n)
    elif hkind == "dense_lowrank":
        # fewer samples than features: H = X^T X / T is rank deficient (lambda_min = 0 up to rounding), the case
        # sigma_reg exists for (alg.py:59-64)
        X = torch.randn(n // 2, n, generator=g) * (0.25 + torch.rand(n, generator=g))[None, :]
        H = X.T @ X / (n // 2)
    elif hkind == "zero_entry":
        h = 0.5 + torch.rand(n, generator=g)
        h[7] = 0.0
        H = torch.diag(h)
    elif hkind == "none":
        H = None
    else:
        raise ValueError(hkind)
    return W, H


CASES = {
    # name: (m, n, seed, hkind, params-kwargs, scale_W)
    "q2_lr16": (96, 128, 1001, "diag",
                dict(Q_bits=2, L_bits=16, R_bits=16, rank=8, iters=4, lplr_iters=5,
                     update_order=["Q", "LR"]), True),
    "q2_lr4": (96, 128, 1002, "diag",
               dict(Q_bits=2, L_bits=4, R_bits=4, rank=8, iters=4, lplr_iters=3,
                    update_order=["Q", "LR"]), True),
    "q4_lr4_lrfirst_heavy": (128, 96, 1003, "heavy",
                             dict(Q_bits=4, L_bits=4, R_bits=4, rank=12, iters=3, lplr_iters=4,
                                  update_order=["LR", "Q"]), True),
    "q2_lr4_dense": (96, 128, 1004, "dense",
                     dict(Q_bits=2, L_bits=4, R_bits=4, rank=8, iters=3, lplr_iters=3,
                          update_order=["Q", "LR"]), True),
    "q4_lr16_dense": (64, 96, 1005, "dense",
                      dict(Q_bits=4, L_bits=16, R_bits=16, rank=8, iters=3,
                           update_order=["Q", "LR"]), True),
    "q2_lr16_identity_nonaware": (96, 128, 1006, "none",
                                  dict(Q_bits=2, L_bits=16, R_bits=16, rank=8, iters=3,
                                       activation_aware_LR=False, update_order=["Q", "LR"]), True),
    "q4_lr4_nonaware_diag": (96, 128, 1007, "diag",
                             dict(Q_bits=4, L_bits=4, R_bits=4, rank=8, iters=3, lplr_iters=3,
                                  activation_aware_LR=False, update_order=["Q", "LR"]), True),
    "q2_lr16_noscale": (96, 128, 1008, "diag",
                        dict(Q_bits=2, L_bits=16, R_bits=16, rank=8, iters=3,
                             update_order=["Q", "LR"]), False),
    "q2_lr16_randsvd": (192, 256, 1009, "diag",
                        dict(Q_bits=2, L_bits=16, R_bits=16, rank=16, iters=3, rand_svd=True,
                             update_order=["Q", "LR"]), True),
    "q4_lr8_sigma_reg": (96, 128, 1010, "zero_entry",
                         dict(Q_bits=4, L_bits=8, R_bits=8, rank=8, iters=3, lplr_iters=2,
                              sigma_reg=1e-3, update_order=["Q", "LR"]), True),
    "q_only": (96, 128, 1011, "diag",
               dict(Q_bits=4, compute_low_rank_factors=False, rank=8, iters=2,
                    update_order=["Q"]), True),
    "lr_only": (96, 128, 1012, "diag",
                dict(L_bits=16, R_bits=16, compute_quantized_component=False, rank=8, iters=2,
                     update_order=["LR"]), True),
    "q4_lr8_dense_sigma_reg": (96, 128, 1015, "dense_lowrank",
                               dict(Q_bits=4, L_bits=8, R_bits=8, rank=8, iters=3, lplr_iters=2,
                                    sigma_reg=1e-3, update_order=["Q", "LR"]), True),
    "q2_lr16_dense_sigma_reg": (128, 160, 1016, "dense_lowrank",
                                dict(Q_bits=2, L_bits=16, R_bits=16, rank=12, iters=3,
                                     sigma_reg=5e-2, update_order=["Q", "LR"]), True),
    "q2_lr16_mid": (384, 512, 1013, "diag",
                    dict(Q_bits=2, L_bits=16, R_bits=16, rank=32, iters=5,
                         update_order=["Q", "LR"]), True),
    "q2_lr4_mid": (512, 384, 1014, "heavy",
                   dict(Q_bits=2, L_bits=4, R_bits=4, rank=32, iters=4, lplr_iters=5,
                        update_order=["Q", "LR"]), True),
}


def caldera_cases(only=None):
    for name, (m, n, seed, hkind, kw, scale_W) in CASES.items():
        if only and name not in only:
            continue
        W, H = make_inputs(m, n, seed, hkind)
        params = CalderaParams(quant_factory_Q=QuantizerFactory(method="uniform", block_size=64),
                               quant_factory_LR=QuantizerFactory(method="uniform", block_size=64),
                               **kw)
        torch.manual_seed(42)
        d = caldera(params, W, H, device="cpu", use_tqdm=False, scale_W=scale_W)
        out = {"W": t2n(W), "scale_W": np.array(scale_W),
               "params": np.array(json.dumps(kw)),
               "global_scale": np.array(d.global_scale, dtype=np.float64),
               "L": t2n(d.L), "R": t2n(d.R)}
        if m * n <= 20000:
            out["Q"] = t2n(d.Q)
            out["W_scaled"] = t2n(d.W)
        if H is not None:
            hd = torch.diagonal(H)
            if torch.equal(torch.diag(hd), H):
                out["h"] = t2n(hd)
            else:
                out["H"] = t2n(H)
        for k, v in d.errors.items():
            out[f"errors_{k}"] = np.array(v, dtype=np.float64)
        for k in ("Q_idxs", "L_idxs", "R_idxs", "Q_scale", "L_scale", "R_scale"):
            v = getattr(d, k)
            if torch.is_tensor(v):
                out[k] = t2n(v)
        np.savez_compressed(os.path.join(OUT, f"caldera_{name}.npz"), **out)
        errs = {k: [round(e, 6) for e in v] for k, v in d.errors.items()}
        print(name, "global_scale", d.global_scale, errs)


if __name__ == "__main__":
    torch.set_num_threads(8)
    import sys
    if len(sys.argv) > 1:                 # regenerate the named caldera cases only
        caldera_cases(set(sys.argv[1:]))
        raise SystemExit(0)
    quantizer_cases()
    caldera_cases()
    with open(os.path.join(OUT, "meta.json"), "w") as f:
        json.dump({"torch": torch.__version__, "numpy": np.__version__,
                   "reference": "genglongling/EE274_ConvexCaldera_LLM_quantization",
                   "generator": "tests/golden/make_golden.py"}, f, indent=1)
