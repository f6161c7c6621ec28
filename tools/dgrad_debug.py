import os, sys
import numpy as np, torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_autograd_ops as T
from uncltmo_b200 import autograd as A
for (ci, co, h) in [(512, 64, 57), (512, 64, 24), (256, 64, 57), (512, 128, 57), (512, 32, 57)]:
    x = T.rnd(2, ci, h, h + 2, seed=1).to(torch.bfloat16).float()
    w = T.rnd(ci, co, 3, 3, seed=2, scale=(9 * ci) ** -0.5).to(torch.bfloat16).float()
    b = T.rnd(co, seed=3, scale=0.1)
    xr, wr, br = T.leaf(x), T.leaf(w), T.leaf(b)
    yr = F.relu(F.conv_transpose2d(xr, wr, br))
    g = T.rnd(*yr.shape, seed=4).to(torch.bfloat16).float()
    yr.backward(g.double())
    xb, wc, bc = T.leaf(T.to_blocked(x), "cuda"), T.leaf(w, "cuda"), T.leaf(b, "cuda")
    y = A.Conv3x3.apply(xb, wc, bc, True, True, True)
    y.backward(T.to_blocked(g).cuda())
    d = (T.from_blocked(xb.grad).double().cpu() - xr.grad).abs()
    bad = (d > 1e-4 * xr.grad.abs().max()).nonzero()
    print(ci, co, h, "rel", T.rel(T.from_blocked(xb.grad), xr.grad), "nbad", bad.shape[0], "of", d.numel())
    if bad.shape[0]:
        print("  n", bad[:, 0].unique().tolist(), "c range", bad[:, 1].min().item(), bad[:, 1].max().item(), "y", bad[:, 2].unique().tolist()[:12], "x", bad[:, 3].unique().tolist()[:12])
