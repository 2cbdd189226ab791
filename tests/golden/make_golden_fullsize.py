"""Full-size goldens (BASELINE.json configs 2 and 3) from the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference; ~10 minutes of CPU on 8 threads):

    python tests/golden/make_golden_fullsize.py [c2] [c3] [c3t]

Synthetic layers of SURVEY.md section 8d (W = 0.02 N(0,1), h = 0.5 + U(0,1), per-layer seed), H = diag(h)
passed as a dense matrix like main.py:165, device="cpu", torch.manual_seed(42) before the call.

  c2   4096 x 4096,  rank 128, Q 2-bit, L/R 16-bit, iters 5          seed 1000  -> fullsize_c2.json  (+ _codes.npz)
  c3   11008 x 4096, rank 256, Q 2-bit, L/R 4-bit, iters 5, lplr 5   seed 1004  -> fullsize_c3.json  (+ _codes.npz)
  c3t  4096 x 11008, same parameters                                 seed 1005  -> fullsize_c3t.json (+ _codes.npz)

The JSON keeps scalars (error trajectory, global_scale, scales, best step) and SHA-256 digests of the
integer codes: `q_idxs_iter0_sha256` (the first Q update quantises W / global_scale itself, so the B200
path must reproduce it bit for bit when the reference's global_scale is injected) and `q_idxs_best_sha256`.
The .npz keeps the best iterate's Q codes bit-packed (MSB-first, offset binary -- the packed wire format),
which compress to a few hundred KiB because the 2-bit grid with one scale per tensor is mostly zeros; the
GPU tests and bench.py report the exact-match fraction against them.  Valid for the torch build named in
the file.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

REF = "/root/reference/rank-constrained-regression-main"
sys.path.insert(0, REF)
from src.caldera.utils.quantization import QuantizerFactory  # noqa: E402
from src.caldera.utils.dataclasses import CalderaParams  # noqa: E402
from src.caldera.decomposition.alg import caldera  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "c2": dict(m=4096, n=4096, seed=1000, rank=128, lbits=16, iters=5,
               config="4096x4096 rank 128 Q2 L/R16 iters 5 order Q,LR aware exact-SVD, synthetic layer seed 1000"),
    "c3": dict(m=11008, n=4096, seed=1004, rank=256, lbits=4, iters=5,
               config="11008x4096 rank 256 Q2 L/R4 iters 5 lplr 5 order Q,LR aware exact-SVD, synthetic seed 1004"),
    "c3t": dict(m=4096, n=11008, seed=1005, rank=256, lbits=4, iters=5,
                config="4096x11008 rank 256 Q2 L/R4 iters 5 lplr 5 order Q,LR aware exact-SVD, synthetic seed 1005"),
}


def pack2(codes: np.ndarray) -> np.ndarray:
    """2-bit codes {-1,0,1} -> offset binary {0,1,2}, four per byte, first element in the top bits."""
    u = (codes.astype(np.int16).reshape(-1) + 1).astype(np.uint8)
    u = u.reshape(-1, 4)
    return (u[:, 0] << 6 | u[:, 1] << 4 | u[:, 2] << 2 | u[:, 3]).astype(np.uint8)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run(name: str) -> None:
    c = CASES[name]
    m, n = c["m"], c["n"]
    g = torch.Generator().manual_seed(c["seed"])
    W = 0.02 * torch.randn(m, n, generator=g, dtype=torch.float32)
    h = 0.5 + torch.rand(n, generator=g, dtype=torch.float32)
    qf = QuantizerFactory(method="uniform", block_size=64)
    params = CalderaParams(compute_quantized_component=True, compute_low_rank_factors=True, Q_bits=2,
                           L_bits=c["lbits"], R_bits=c["lbits"], rank=c["rank"], iters=c["iters"], lplr_iters=5,
                           activation_aware_LR=True, update_order=["Q", "LR"], quant_factory_Q=qf,
                           quant_factory_LR=qf, rand_svd=False, sigma_reg=0.0)
    torch.manual_seed(42)
    t0 = time.time()
    dec = caldera(params, W, torch.diag(h), device="cpu", use_tqdm=False, scale_W=True)
    dt = time.time() - t0
    errs = {k: [float(e) for e in v] for k, v in dec.errors.items()}
    flat = [e for pair in zip(errs["Q"], errs["LR"]) for e in pair]
    best = min(range(1, len(flat)), key=lambda i: flat[i])   # strict arg-min once both components were updated
    # iterate 0: the first Q update quantises W / global_scale with one scale per tensor (alg.py:247, 262)
    q0 = QuantizerFactory(method="uniform", block_size=m * n).get_quantizer(2, "cpu")
    codes0, scale0, _ = q0.quantize_block(W / dec.global_scale)
    q_best = dec.Q_idxs.numpy().reshape(-1)
    out = {"config": c["config"], "errors": errs, "global_scale": float(dec.global_scale),
           "Q_scale": float(dec.Q_scale.reshape(-1)[0]), "Q_scale_iter0": float(scale0.reshape(-1)[0]),
           "best_step": best, "best_error": flat[best],
           "q_idxs_iter0_sha256": sha(codes0.numpy().reshape(-1)), "q_idxs_best_sha256": sha(q_best),
           "q_nonzero_fraction_best": float(np.count_nonzero(q_best)) / q_best.size,
           "torch": torch.__version__, "reference_cpu_seconds": dt, "threads": torch.get_num_threads()}
    if c["lbits"] < 16:
        out["L_scale"] = float(dec.L_scale.reshape(-1)[0])
        out["R_scale"] = float(dec.R_scale.reshape(-1)[0])
        out["l_idxs_sha256"] = sha(dec.L_idxs.numpy().reshape(-1))
        out["r_idxs_sha256"] = sha(dec.R_idxs.numpy().reshape(-1))
    with open(os.path.join(OUT, f"fullsize_{name}.json"), "w") as f:
        json.dump(out, f, indent=1)
    np.savez_compressed(os.path.join(OUT, f"fullsize_{name}_codes.npz"), q_packed_best=pack2(q_best))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    for nm in (sys.argv[1:] or list(CASES)):
        run(nm)
