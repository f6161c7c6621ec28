"""Achieved HBM bandwidth of the memory-bound kernels at a size where they are not launch-latency bound (4K frame / 220
tiles, 64-image loss batches).  Algorithmic bytes (compulsory reads + writes) / device time, against MEASURED_PEAKS.json."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from uncltmo_b200 import _lib, synth
from uncltmo_b200.frame import FramePipeline
from uncltmo_b200.generator import UNet
from uncltmo_b200.struct_loss import StructLoss
from uncltmo_b200 import losses
from uncltmo_b200.weights import make_generator_state_dict

G_ARGS = (1, 1, "sigmoid", 4, 4, "square_and_square_root", 32, 0, "unet", 0, 0, "none", "none", "relu", 1, "replicate", 2)
peak = 6453.0
p = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
if os.path.exists(p):
    d = json.load(open(p))
    for k in ("hbm_gbs", "hbm_gbps"):
        if k in d:
            peak = float(d[k])


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


rows = []


def report(name, nbytes, ms):
    gbps = nbytes / ms / 1e6
    rows.append((name, nbytes / 1e6, ms * 1e3, gbps, gbps / peak))
    print("%-44s %9.1f MB %9.1f us %8.0f GB/s  %5.1f %% of %.0f" % (name, nbytes / 1e6, ms * 1e3, gbps, 100 * gbps / peak, peak), flush=True)


torch.set_grad_enabled(False)
net = UNet(*G_ARGS, up_mode=0, precision="bf16").cuda().eval()
net.load_state_dict(make_generator_state_dict())
pipe = FramePipeline(net)
H, W = 2160, 3840
rgb = torch.from_numpy(synth.hdr_frame(H, W, seed=0)).cuda()
pl = pipe.plan(H, W, rgb.device)
HW, HW1 = H * W, pl.h1 * pl.w1
gray_p, stats = pipe.normalise_pad(rgb, 50.0)
report("frame_normalise_pad (4K, 5 launches)", HW * 12 * 2 + HW1 * 4, timed(lambda: pipe.normalise_pad(rgb, 50.0)))
tiles = pipe.gather_tiles(gray_p, pl)
report("tiles_gather (220 tiles)", pl.ntiles * 65536 * 8, timed(lambda: pipe.gather_tiles(gray_p, pl)))
report("tiles_blend (220 tiles)", pl.ntiles * 65536 * 4 + HW1 * 4, timed(lambda: pipe.blend(tiles, pl)))
fake_p = pipe.blend(tiles, pl)
report("percentile_pair (4K plane, 3 passes)", 3 * HW1 * 4, timed(lambda: pipe.percentiles(fake_p, 0.5, 99.5)))
report("frame_postprocess (+ its percentile pair)", 3 * HW1 * 4 + HW1 * 4 + HW * 24, timed(lambda: pipe.postprocess(fake_p, rgb, stats, pl)))
col = pipe.postprocess(fake_p, rgb, stats, pl)
report("frame_to_u8 (+ its percentile pair)", 3 * HW * 12 + HW * 15, timed(lambda: pipe.to_uint8(col)))
# the fused cooperative stage kernels at the size they are built for (1080p: every CTA's order keys fit its shared memory)
H2, W2 = 1080, 1920
rgb2 = torch.from_numpy(synth.hdr_frame(H2, W2, seed=1)).cuda()
pl2 = pipe.plan(H2, W2, rgb2.device)
assert pipe.fused_ok(pl2)
hw2, hw12 = H2 * W2, pl2.h1 * pl2.w1
tiles2, stats2 = pipe.normalise_tiles(rgb2, 50.0, pl2)
report("frame_normalise_tiles (1080p, 1 cooperative launch)", hw2 * 24 + pl2.ntiles * 65536 * 4, timed(lambda: pipe.normalise_tiles(rgb2, 50.0, pl2)))
fake2, pct2 = pipe.blend_percentiles(tiles2, pl2)
report("frame_blend_percentiles (1080p, 1 launch)", pl2.ntiles * 65536 * 4 + hw12 * 4, timed(lambda: pipe.blend_percentiles(tiles2, pl2)))
report("frame_post_u8 (1080p, 1 launch)", hw12 * 4 + hw2 * 15, timed(lambda: pipe.post_uint8(fake2, pct2, rgb2, stats2, pl2)))
# generator-side memory-bound kernels at 220 tiles
n = 220
x = torch.rand(n, 1, 256, 256, device="cuda")
a0 = torch.empty((n, 4, 254, 254, 8), device="cuda", dtype=torch.bfloat16)
P = net.packed()
report("conv_first (220 tiles, bf16 out)", n * (65536 * 4 + 254 * 254 * 64),
       timed(lambda: _lib.call("uncl_conv_first", x, P["inc0"][0], P["inc0"][1], a0, a0.stride(0), n, 256, 256, 32, _lib.ACT_RELU, _lib.BF16)))
cur = torch.randn((n, 4, 252, 252, 8), device="cuda").bfloat16()
pooled = torch.empty((n, 4, 126, 126, 8), device="cuda", dtype=torch.bfloat16)
report("maxpool2 (220 tiles, 32 ch, 252^2)", n * 32 * (252 * 252 + 126 * 126) * 2,
       timed(lambda: _lib.call("uncl_maxpool2", cur, cur.stride(0), None, 0, 0, pooled, pooled.stride(0), n, 32, 252, 252, _lib.BF16)))
# losses at a 64-image batch
b = 64
fake = torch.rand(b, 1, 256, 256, device="cuda")
hdr = torch.rand(b, 1, 256, 256, device="cuda")
sl = StructLoss([1.0, 1.0, 1.0])
report("StructLoss forward (64 images, 3 levels)", b * 65536 * 8 * (1 + 0.25 + 0.0625) + b * 65536 * 8 * 0.3125, timed(lambda: sl(fake, None, hdr, [1.0, 1.0, 1.0])))
fa = torch.randn(16, 32, 256, 256, device="cuda")
report("nce forward (image features 16x32x256^2, pos+neg broadcast)", 16 * 32 * 65536 * 4 * 2 + 2 * 32 * 65536 * 4,
       timed(lambda: losses.nce_from_indices(fa, 3, 5, "InfoNCE", 1.0, 1e-2)))
report("L_TV forward (64 images)", b * 65536 * 4, timed(lambda: losses.L_TV()(fake)))
report("plane_mean_contrast (64 planes)", b * 65536 * 4, timed(lambda: losses.plane_mean_contrast(fake)))
with open(os.path.join(os.path.dirname(__file__), "..", "gpurun_out", "hbm_kernels.txt"), "w") as f:
    f.write("kernel | algorithmic MB | us | GB/s | fraction of %.0f GB/s\n" % peak)
    for r in rows:
        f.write("%s | %.1f | %.1f | %.0f | %.3f\n" % r)
