"""Diagnostic: device memory held after each eager training step / after a capture / after replays."""
import gc, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from uncltmo_b200 import synth
from uncltmo_b200.discriminator import SimpleDiscriminator
from uncltmo_b200.generator import UNet
from uncltmo_b200.optim import FlatAdam
from uncltmo_b200.trainer import GanTrainerStep
from uncltmo_b200.weights import make_discriminator_state_dict, make_generator_state_dict
G_ARGS = (1, 1, "sigmoid", 4, 4, "square_and_square_root", 32, 0, "unet", 0, 0, "none", "none", "relu", 1, "replicate", 2)
GB = lambda: torch.cuda.memory_allocated() / 2**30
netG = UNet(*G_ARGS, up_mode=0, precision="bf16").cuda().train(); netG.load_state_dict(make_generator_state_dict())
netD = SimpleDiscriminator(256, 1, 16, "none", "none", 0, 0).cuda().train(); netD.load_state_dict(make_discriminator_state_dict())
tr = GanTrainerStep(netG, netD, FlatAdam(netG, lr=1e-5, betas=(0.5, 0.999)),
                    torch.optim.Adam(netD.parameters(), lr=1.5e-5, betas=(0.5, 0.999), capturable=True, fused=True))
B = 8
hdr = torch.from_numpy(synth.normalised_batch(2 * B, seed=4)).reshape(B, 2, 1, 256, 256).cuda()
pos = torch.from_numpy(synth.ldr_batch(2 * B, seed=5)).reshape(B, 2, 1, 256, 256).cuda()
neg = torch.from_numpy(synth.ldr_batch(2 * B, seed=6)).reshape(B, 2, 1, 256, 256).cuda()
print("start %.2f GB" % GB())
for flag in (True, False):
    tr.overlap_g_forward = flag
    for i in range(4):
        tr.step(hdr, None, pos, neg, 0)
        torch.cuda.synchronize()
        print("eager step %d (G forward next to the D step: %s): %.2f GB" % (i, flag, GB()), flush=True)
gc.collect(); torch.cuda.empty_cache(); print("after gc %.2f GB" % GB())
tr.overlap_g_forward = True
tr.capture(hdr, None, pos, neg, 0, warmup=1)
print("after capture %.2f GB" % GB())
for i in range(3):
    tr.replay(hdr, None, pos, neg, 0)
torch.cuda.synchronize(); print("after replays %.2f GB" % GB())
del tr
gc.collect(); torch.cuda.empty_cache(); print("trainer deleted %.2f GB" % GB())
del netG, netD
gc.collect(); torch.cuda.empty_cache(); print("networks deleted %.2f GB" % GB())
