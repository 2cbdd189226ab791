"""How far the CUDA path is from each stored reference run (tests/golden/caldera_*.npz): max relative deviation of the
error trajectory and of the best error.  Used to set the tolerances in tests/test_gpu_caldera.py::test_golden."""
import glob
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from test_gpu_caldera import _load, _params  # noqa: E402
from src.caldera.decomposition.alg import caldera  # noqa: E402

for path in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "caldera_*.npz"))):
    z, kw, H = _load(path)
    p = _params(kw)
    scale_W = bool(z["scale_W"])
    gs = float(z["global_scale"]) if scale_W else None
    d = caldera(p, torch.from_numpy(z["W"]), H, device="cuda", use_tqdm=False, scale_W=scale_W, global_scale=gs)
    dev = 0.0
    for k in p.update_order:
        ref, got = z[f"errors_{k}"], np.array(d.errors[k])
        dev = max(dev, float(np.max(np.abs(got - ref) / np.abs(ref))))
    order, kk = p.update_order, len(p.update_order)
    seq_ref = [float(z[f"errors_{order[s % kk]}"][s // kk]) for s in range(p.iters * kk)]
    seq_got = [d.errors[order[s % kk]][s // kk] for s in range(p.iters * kk)]
    rb, gb = min(seq_ref[kk - 1:]), min(seq_got[kk - 1:])
    quantised = p.compute_low_rank_factors and (p.L_bits < 16 or p.R_bits < 16)
    print(f"{os.path.basename(path):45s} quantised_factors={quantised!s:5s} rand_svd={p.rand_svd!s:5s} shape={z['W'].shape} "
          f"trajectory_dev={dev:.2e} best_rel_diff={(gb - rb) / rb:+.2e}")
