import sys, time; sys.path.insert(0, "/root/repo")
import torch
from src.caldera.utils.dataclasses import CalderaParams
from src.caldera.utils.quantization import QuantizerFactory
from src.caldera.decomposition.alg import caldera
qf = QuantizerFactory(method="uniform", block_size=64)
for (m, n) in ((8192, 28672), (28672, 8192)):
    g = torch.Generator(device="cuda").manual_seed(m)
    W = 0.02 * torch.randn(m, n, generator=g, device="cuda")
    h = 0.5 + torch.rand(n, generator=g, device="cuda")
    p = CalderaParams(Q_bits=2, L_bits=16, R_bits=16, rank=128, iters=2, update_order=["Q", "LR"], quant_factory_Q=qf, quant_factory_LR=qf)
    torch.cuda.synchronize(); t0 = time.time()
    d = caldera(p, W, h, device="cuda", use_tqdm=False, W_copy="none")
    torch.cuda.synchronize(); dt = time.time() - t0
    E = (d.Q + d.L @ d.R) * d.global_scale - W
    err = float(((E * E) * h).sum().sqrt() / ((W * W) * h).sum().sqrt())
    flat = [e for pair in zip(d.errors["Q"], d.errors["LR"]) for e in pair]
    print(f"{m}x{n}: {dt:.2f} s, errors {d.errors}, recomputed best {err:.6f} vs reported {flat[d.best_step]:.6f}, stats {d.device_stats}, peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
    del W, d, E
    torch.cuda.empty_cache()
