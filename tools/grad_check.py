import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_inputs as gi, oracle
from uncltmo_b200.generator import UNet
from uncltmo_b200.weights import make_generator_state_dict
G_ARGS = (1, 1, "sigmoid", 4, 4, "square_and_square_root", 32, 0, "unet", 0, 0, "none", "none", "relu", 1, "replicate", 2)
def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()
sd = make_generator_state_dict(); x = gi.generator_input()[:1]
rng = np.random.default_rng(3)
po = torch.from_numpy(rng.standard_normal((1, 1, 256, 256)).astype(np.float32))
pf = torch.from_numpy((rng.standard_normal((1, 32, 256, 256)) * 0.01).astype(np.float32))
def ograds(dtype, wo, wf):
    params = {k: v.clone().to(dtype).requires_grad_(k != "gcn.module.0.0.relative_pos") for k, v in sd.items()}
    out, feats = oracle.unet_forward(params, x.to(dtype))
    ((out * po.to(dtype)).sum() * wo + (feats * pf.to(dtype)).sum() * wf).backward()
    return {k: v.grad for k, v in params.items() if v.grad is not None}
for wo, wf in ((1.0, 0.0), (0.0, 1.0), (1.0, 1.0)):
    r32 = ograds(torch.float32, wo, wf); r64 = ograds(torch.float64, wo, wf)
    net = UNet(*G_ARGS, up_mode=0, precision="fp32").cuda().train(); net.load_state_dict(sd); net.drop_path_prob = 0.0
    out, feats = net(x.cuda())
    ((out * po.cuda()).sum() * wo + (feats * pf.cuda()).sum() * wf).backward()
    got = dict(net.named_parameters())
    print("=== wo=%g wf=%g" % (wo, wf))
    for k in r64:
        print("%-45s cuda-vs-f64 %.2e  cpu32-vs-f64 %.2e  cuda-vs-cpu32 %.2e" % (k, rel(got[k].grad, r64[k]), rel(r32[k], r64[k]), rel(got[k].grad, r32[k])))
