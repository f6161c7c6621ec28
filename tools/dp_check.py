"""Data-parallel parity of the training step (torchrun, one process per GPU): the 16-image batch of the reference-trainer
fixture split over the ranks must reproduce the single-process losses (sum of the per-rank shares) and gradient norms
(after the all-reduce) - including infoNCE2 / pseudo_label_loss, whose arg-max / arg-min run over the GLOBAL batch.

    torchrun --nproc-per-node 2 tools/dp_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_inputs as gi  # noqa: E402
from uncltmo_b200.discriminator import SimpleDiscriminator  # noqa: E402
from uncltmo_b200.generator import UNet  # noqa: E402
from uncltmo_b200.trainer import GanTrainerStep  # noqa: E402
from uncltmo_b200.weights import make_discriminator_state_dict, make_generator_state_dict  # noqa: E402

G_ARGS = (1, 1, "sigmoid", 4, 4, "square_and_square_root", 32, 0, "unet", 0, 0, "none", "none", "relu", 1, "replicate", 2)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
golden = np.load(os.path.join(ROOT, "tests", "golden", "reference_train_outputs.npz"))
ok = True
for epoch in (0, 7, 10):
    hdr, pos, neg = gi.train_batch(16)
    per = hdr.shape[0] // world
    sl = slice(rank * per, (rank + 1) * per)
    netG = UNet(*G_ARGS, up_mode=0, precision="bf16").to(dev).train()
    netG.load_state_dict(make_generator_state_dict())
    netG.drop_path_prob = 0.0
    netD = SimpleDiscriminator(256, 1, 16, "none", "none", 0, 0).to(dev).train()
    netD.load_state_dict(make_discriminator_state_dict())
    tr = GanTrainerStep(netG, netD, torch.optim.SGD([p for p in netG.parameters() if p.requires_grad], lr=0.0),
                        torch.optim.SGD(netD.parameters(), lr=0.0))
    eg, es = tr.step(hdr[sl].to(dev), None, pos[sl].to(dev), neg[sl].to(dev), epoch)
    shares = torch.stack([eg.detach(), es.detach()])
    dist.all_reduce(shares)
    tag = "ref/img/e%d/b16/" % epoch
    # errD / the contrastive part use gathered logits (identical on every rank); the per-sample terms are shares
    want_g, want_s = float(golden[tag + "errG_d"]), float(golden[tag + "errG_struct"])
    # errG_d = gathered contrastive term (counted on every rank) + per-sample shares: compare through the struct loss
    # (pure share) and the gradient norms; errD is identical on all ranks
    rel_s = abs(shares[1].item() - want_s) / want_s
    rel_d = abs(tr.errD.item() - float(golden[tag + "errD"])) / float(golden[tag + "errD"])
    gn = {}
    for k in ("up_path.3.conv.conv1.weight", "up_path.0.conv.conv.weight", "outc.conv.weight", "gcn.module.0.1.fc2.0.weight"):
        want = golden["obf/img/e%d/b16/gG/%s" % (epoch, k)][0]
        got = dict(netG.named_parameters())[k].grad.double().norm().item()
        gn[k] = abs(got - want) / want
    if rank == 0:
        print("dp%d epoch %d: errD rel %.1e, errG_struct (sum of shares) rel %.1e, gradient-norm deviations after all-reduce %s"
              % (world, epoch, rel_d, rel_s, {k: "%.1e" % v for k, v in gn.items()}), flush=True)
    ok = ok and rel_s <= 1e-3 and rel_d <= 1e-3 and max(gn.values()) <= 5e-2
dist.barrier()
if rank == 0:
    print("dp_check", "OK" if ok else "FAILED")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
