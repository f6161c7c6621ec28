"""Frames/s of the 1080p frame path with the skip operators materialised vs fused into the consuming convs, measured over
long back-to-back runs (the GPU sits at its power cap under this load: less DRAM traffic may buy clock)."""
import os, sys
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
N = int(sys.argv[1]) if len(sys.argv) > 1 else 600
with torch.no_grad():
    for r in range(3):
        for fs in (False, True):
            net.fused_skip = fs
            for i in range(10):
                pipe.tonemap(frames[i % 3], 50.0, uint8=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(N):
                pipe.tonemap(frames[i % 3], 50.0, uint8=True)
            e1.record()
            torch.cuda.synchronize()
            print("round %d  fused_skip %-5s  %d frames  %.1f frames/s  %.3f ms/frame" % (r, fs, N, N * 1e3 / e0.elapsed_time(e1), e0.elapsed_time(e1) / N), flush=True)
