"""Diagnostic: eager vs eager vs CUDA-graph replay trajectories (GPU box)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from uncltmo_b200 import synth
from uncltmo_b200.discriminator import SimpleDiscriminator
from uncltmo_b200.generator import UNet
from uncltmo_b200.trainer import GanTrainerStep
from uncltmo_b200.weights import make_discriminator_state_dict, make_generator_state_dict
G_ARGS = (1, 1, "sigmoid", 4, 4, "square_and_square_root", 32, 0, "unet", 0, 0, "none", "none", "relu", 1, "replicate", 2)
PREC = sys.argv[1] if len(sys.argv) > 1 else "bf16"
hdr = torch.from_numpy(synth.normalised_batch(4, seed=4)).reshape(2, 2, 1, 256, 256).cuda()
pos = torch.from_numpy(synth.ldr_batch(4, seed=5)).reshape(2, 2, 1, 256, 256).cuda()
neg = torch.from_numpy(synth.ldr_batch(4, seed=6)).reshape(2, 2, 1, 256, 256).cuda()
hdr2 = torch.from_numpy(synth.normalised_batch(4, seed=14)).reshape(2, 2, 1, 256, 256).cuda()


def make():
    netG = UNet(*G_ARGS, up_mode=0, precision=PREC).cuda().train()
    netG.load_state_dict(make_generator_state_dict())
    netG.drop_path_prob = 0.0
    netD = SimpleDiscriminator(256, 1, 16, "none", "none", 0, 0).cuda().train()
    netD.load_state_dict(make_discriminator_state_dict())
    optG = torch.optim.Adam([p for p in netG.parameters() if p.requires_grad], lr=1e-4, betas=(0.5, 0.999), capturable=True)
    optD = torch.optim.Adam(netD.parameters(), lr=1e-4, betas=(0.5, 0.999), capturable=True)
    return netG, netD, GanTrainerStep(netG, netD, optG, optD)


def eager(n):
    g, d, tr = make()
    for b in (hdr, hdr, hdr2, hdr)[:n]:
        tr.step(b, None, pos, neg, 0)
    return g, d, tr


init = {k: v.cuda() for k, v in make_generator_state_dict().items()}
for n in (1, 2, 3, 4):
    gA, _, trA = eager(n)
    gA2, _, trA2 = eager(n)
    gB, dB, trB = make()
    if n >= 2:
        trB.capture(hdr, None, pos, neg, 0, warmup=2)
        for b in (hdr2, hdr)[:n - 2]:
            trB.replay(b, None, pos, neg, 0)
    else:
        trB.step(hdr, None, pos, neg, 0)
    torch.cuda.synchronize()
    worst_ee = worst_eg = 0.0
    for (k, a), (_, a2), (_, b) in zip(gA.named_parameters(), gA2.named_parameters(), gB.named_parameters()):
        moved = (a.detach() - init[k]).norm().item() + 1e-12
        worst_ee = max(worst_ee, (a - a2).norm().item() / moved)
        worst_eg = max(worst_eg, (a - b).norm().item() / moved)
    print("steps %d: eager-vs-eager %.3e   eager-vs-graph %.3e   errD %.6f %.6f %s" % (
        n, worst_ee, worst_eg, trA.errD.item(), trA2.errD.item(), trB.errD.item() if trB.errD is not None else None), flush=True)
