"""Times the pixel-shuffle up-conv GEMM (uncl_convT2x2_tc) at the generator's four 60-tile shapes and the percentile
pair at the two 1080p plane sizes; checks the up-conv against torch's conv_transpose2d."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import torch.nn.functional as F
from uncltmo_b200 import _lib, packing
from uncltmo_b200.frame import FramePipeline


def timed(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


g = torch.Generator(device="cuda").manual_seed(0)
for C, H, H2 in ((256, 12, 24), (128, 28, 57), (64, 61, 122), (32, 126, 252)):
    n = 60
    x = torch.randn((n, C // 8, H, H, 8), device="cuda", generator=g).bfloat16()
    w = (torch.randn((C, C, 2, 2), device="cuda", generator=g) / C ** 0.5).bfloat16().float()
    b = torch.randn(C, device="cuda", generator=g) * 0.1
    out = torch.zeros((n, C // 8, H2, H2, 8), device="cuda", dtype=torch.bfloat16)
    wp = packing.convT2x2_tc(w)
    fn = lambda: _lib.call("uncl_convT2x2_tc", x, x.stride(0), wp, b, out, out.stride(0), _lib.BF16, n, C, H, H, H2, H2)
    us = timed(fn)
    xn = x[:2].float().permute(0, 1, 4, 2, 3).reshape(2, C, H, H)
    ref = F.conv_transpose2d(xn, w, b, stride=2)
    if H2 != 2 * H:
        ref = F.pad(ref, (0, 1, 0, 1), mode="replicate")
    got = out[:2].float().permute(0, 1, 4, 2, 3).reshape(2, C, H2, H2)
    rel = ((got - ref).norm() / ref.norm()).item()
    mb = (x.numel() + out.numel()) * 2 / 1e6
    print("convT2x2_tc C=%3d %3d->%3d: %6.1f us  %6.0f GB/s  rel err %.2e" % (C, H, H2, us, mb / us * 1e3, rel), flush=True)

x = torch.rand((60, 1, 256, 256), device="cuda", generator=g)
wt = torch.randn((32, 1, 3, 3), device="cuda", generator=g) * 0.3
b = torch.randn(32, device="cuda", generator=g) * 0.1
o = torch.empty((60, 4, 254, 254, 8), device="cuda", dtype=torch.bfloat16)
w9, ws = packing.conv_first(wt), packing.conv_first_tc_split(wt)
print("conv_first (CUDA cores): %6.1f us" % timed(lambda: _lib.call("uncl_conv_first", x, w9, b, o, o.stride(0), 60, 256, 256, 32, 1, _lib.BF16)))
print("conv_first_tc          : %6.1f us" % timed(lambda: _lib.call("uncl_conv_first_tc", x, ws, b, o, o.stride(0), 60, 256, 256, 32, 1)))

pipe = FramePipeline(None)
for shape in ((1088, 1936), (1080, 1920, 3)):
    d = torch.rand(shape, device="cuda", generator=g)
    us = timed(lambda: pipe.percentiles(d, 0.5, 99.5))
    ref = torch.quantile(d.flatten()[:16000000].double(), torch.tensor([0.005, 0.995], device="cuda", dtype=torch.float64))
    got = pipe.percentiles(d, 0.5, 99.5)
    print("percentile_pair %s: %6.1f us  got %s  torch %s" % (shape, us, [round(v, 7) for v in got.tolist()], [round(v, 7) for v in ref.tolist()]))
