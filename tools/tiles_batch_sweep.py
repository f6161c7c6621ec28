"""Generator time per 256x256 tile as a function of the tiles per call (60 = one 1080p frame)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from uncltmo_b200.generator import UNet
from uncltmo_b200.weights import make_generator_state_dict
G_ARGS = (1, 1, "sigmoid", 4, 4, "square_and_square_root", 32, 0, "unet", 0, 0, "none", "none", "relu", 1, "replicate", 2)
net = UNet(*G_ARGS, up_mode=0, precision="bf16").cuda().eval()
net.load_state_dict(make_generator_state_dict())
with torch.no_grad():
    for rep in range(2):
        for n in (30, 60, 120, 180, 240):
            x = torch.rand(n, 1, 256, 256, device="cuda")
            for _ in range(3):
                net.tonemap_tiles(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = max(4, 1200 // n)
            e0.record()
            for _ in range(iters):
                net.tonemap_tiles(x)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            print("tiles %4d  %.3f ms per call  %.2f us per tile  -> %.1f frames/s of generator time alone" % (n, ms, 1e3 * ms / n, 1e3 / (ms / n * 60)), flush=True)
