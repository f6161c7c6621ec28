"""Workload for compute-sanitizer (memcheck / racecheck / synccheck): one small bf16 frame through the whole frame path
(tensor-core convs with their mbarrier pipelines and TMEM allocation, the cooperative grid-barrier percentile select,
blend, post-process - as the three fused cooperative stage kernels, as the staged path, and with the skip operators fused
into the consuming convs) and one mixed-precision training step on 2 images (tensor-core dgrad / wgrad with fp32 atomics,
discriminator, every loss forward + backward).

    compute-sanitizer --tool memcheck  python tools/sanitize_run.py            > profiles/r2_sanitizer_memcheck.txt
    compute-sanitizer --tool racecheck python tools/sanitize_run.py frame      > profiles/r2_sanitizer_racecheck_frame.txt
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from uncltmo_b200 import synth  # noqa: E402
from uncltmo_b200.discriminator import SimpleDiscriminator  # noqa: E402
from uncltmo_b200.frame import FramePipeline  # noqa: E402
from uncltmo_b200.generator import UNet  # noqa: E402
from uncltmo_b200.trainer import GanTrainerStep  # noqa: E402
from uncltmo_b200.weights import make_discriminator_state_dict, make_generator_state_dict  # noqa: E402

G_ARGS = (1, 1, "sigmoid", 4, 4, "square_and_square_root", 32, 0, "unet", 0, 0, "none", "none", "relu", 1, "replicate", 2)
what = sys.argv[1] if len(sys.argv) > 1 else "all"

if what in ("all", "frame"):
    net = UNet(*G_ARGS, up_mode=0, precision="bf16").cuda().eval()
    net.load_state_dict(make_generator_state_dict())
    rgb = torch.from_numpy(synth.hdr_frame(268, 300, seed=5)).cuda()
    with torch.no_grad():
        pipe = FramePipeline(net)
        u8 = pipe.tonemap(rgb, 50.0, uint8=True)          # fused cooperative stage kernels
        torch.cuda.synchronize()
        print("frame ok:", tuple(u8.shape), int(u8.min()), int(u8.max()))
        pipe.fused = False                                 # the staged 14-launch path
        u8s = pipe.tonemap(rgb, 50.0, uint8=True)
        net.fused_skip = True                              # skip operators built inside the consuming conv (ring program)
        pipe.fused = True
        u8f = pipe.tonemap(rgb, 50.0, uint8=True)
    torch.cuda.synchronize()
    print("staged == fused:", bool(torch.equal(u8, u8s)), " fused-skip max |diff| (8-bit levels):",
          int((u8f.int() - u8.int()).abs().max()))

if what in ("all", "train"):
    netG = UNet(*G_ARGS, up_mode=0, precision="bf16").cuda().train()
    netG.load_state_dict(make_generator_state_dict())
    netD = SimpleDiscriminator(256, 1, 16, "none", "none", 0, 0).cuda().train()
    netD.load_state_dict(make_discriminator_state_dict())
    optG = torch.optim.Adam([p for p in netG.parameters() if p.requires_grad], lr=1e-5, betas=(0.5, 0.999))
    optD = torch.optim.Adam(netD.parameters(), lr=1.5e-5, betas=(0.5, 0.999))
    tr = GanTrainerStep(netG, netD, optG, optD)
    hdr = torch.from_numpy(synth.normalised_batch(2, seed=4)).reshape(1, 2, 1, 256, 256).cuda()
    pos = torch.from_numpy(synth.ldr_batch(2, seed=5)).reshape(1, 2, 1, 256, 256).cuda()
    neg = torch.from_numpy(synth.ldr_batch(2, seed=6)).reshape(1, 2, 1, 256, 256).cuda()
    for epoch in (0, 7, 10):
        eg, es = tr.step(hdr, None, pos, neg, epoch)
        torch.cuda.synchronize()
        print("train step epoch %d ok: errD %.5f errG_d %.5f errG_struct %.5f" % (epoch, tr.errD.item(), eg.item(), es.item()))
