import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_backward as T
sd, x, po, pf = T._problem()
for i in range(4):
    w = sd["up_path.%d.conv.conv.weight" % i]; w[3 * w.shape[0] // 4:] = 0.0
_, ref = T.oracle_grads({k: v.double() for k, v in sd.items()}, x.double(), po.double(), pf.double())
for prec in ("fp32", "bf16"):
    net = T.UNet(*T.G_ARGS, up_mode=0, precision=prec).cuda().train(); net.load_state_dict(sd); net.drop_path_prob = 0.0
    out, feats = net(x.cuda())
    ((out * po.cuda()).sum() + (feats * pf.cuda()).sum()).backward()
    print(prec, " ".join("%s=%.1e" % (k.replace("mpconv.1.", "").replace(".weight", ".w").replace(".bias", ".b"), T.rel(p.grad[: (3 * ref[k].shape[0] // 4 if ("up_path" in k and k.endswith("conv.conv.weight")) else None)], ref[k][: (3 * ref[k].shape[0] // 4 if ("up_path" in k and k.endswith("conv.conv.weight")) else None)])) for k, p in net.named_parameters() if k in ref and k.endswith("weight")))
