"""Frames/s of the 1080p frame path with each subset of the fused stage kernels (device-resident frames, CUDA events)."""
import itertools, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from uncltmo_b200 import synth
from uncltmo_b200.frame import FramePipeline
from uncltmo_b200.generator import UNet
from uncltmo_b200.weights import make_generator_state_dict

G_ARGS = (1, 1, "sigmoid", 4, 4, "square_and_square_root", 32, 0, "unet", 0, 0, "none", "none", "relu", 1, "replicate", 2)
net = UNet(*G_ARGS, up_mode=0, precision="bf16").cuda().eval()
net.load_state_dict(make_generator_state_dict())
pipe = FramePipeline(net)
frames = [torch.from_numpy(synth.hdr_frame(1080, 1920, seed=s)).cuda() for s in range(3)]
with torch.no_grad():
    for r in range(4):
        for sub in itertools.chain.from_iterable(itertools.combinations(("norm", "blend", "post"), k) for k in range(4)):
            pipe.fused_stages = sub
            for i in range(5):
                pipe.tonemap(frames[i % 3], 50.0, uint8=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(30):
                pipe.tonemap(frames[i % 3], 50.0, uint8=True)
            e1.record()
            torch.cuda.synchronize()
            if r:
                print("round %d  fused %-22s %.1f frames/s  %.3f ms/frame" % (r, "+".join(sub) or "-", 30e3 / e0.elapsed_time(e1), e0.elapsed_time(e1) / 30), flush=True)
