"""One eager mixed-precision training step (16 images) between cudaProfilerStart / Stop, after warm-up steps:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file <out> python tools/train_step_profile.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uncltmo_b200 import synth  # noqa: E402
from uncltmo_b200.discriminator import SimpleDiscriminator  # noqa: E402
from uncltmo_b200.generator import UNet  # noqa: E402
from uncltmo_b200.optim import FlatAdam  # noqa: E402
from uncltmo_b200.trainer import GanTrainerStep  # noqa: E402
from uncltmo_b200.weights import make_discriminator_state_dict, make_generator_state_dict  # noqa: E402

G_ARGS = (1, 1, "sigmoid", 4, 4, "square_and_square_root", 32, 0, "unet", 0, 0, "none", "none", "relu", 1, "replicate", 2)
netG = UNet(*G_ARGS, up_mode=0, precision="bf16").cuda().train()
netG.load_state_dict(make_generator_state_dict())
netD = SimpleDiscriminator(256, 1, 16, "none", "none", 0, 0).cuda().train()
netD.load_state_dict(make_discriminator_state_dict())
tr = GanTrainerStep(netG, netD, FlatAdam(netG, lr=1e-5, betas=(0.5, 0.999)),
                    torch.optim.Adam(netD.parameters(), lr=1.5e-5, betas=(0.5, 0.999), capturable=True, fused=True))
B = 8
hdr = torch.from_numpy(synth.normalised_batch(2 * B, seed=4)).reshape(B, 2, 1, 256, 256).cuda()
pos = torch.from_numpy(synth.ldr_batch(2 * B, seed=5)).reshape(B, 2, 1, 256, 256).cuda()
neg = torch.from_numpy(synth.ldr_batch(2 * B, seed=6)).reshape(B, 2, 1, 256, 256).cuda()
epoch = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for _ in range(3):
    tr.step(hdr, None, pos, neg, epoch)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step(hdr, None, pos, neg, epoch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step at epoch", epoch)
