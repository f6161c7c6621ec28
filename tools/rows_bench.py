"""Row kernel (conv_tc_rows.cu) against the kernels it replaces, per layer of the 1080p frame path, CUDA events.
usage: python tools/rows_bench.py [tiles=240] [reps=5]"""
import sys

import torch

from uncltmo_b200 import _lib, packing

n = int(sys.argv[1]) if len(sys.argv) > 1 else 240
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(1)
# name, logical C_in, H = W, pad, fused skip operators, fused out conv, C_out
LAYERS = [("inc.conv1", 32, 254, 0, False, False, 32), ("down0.conv0", 32, 126, 0, False, False, 64),
          ("down0.conv1", 64, 124, 0, False, False, 64), ("up2.conv0", 256, 122, 2, True, False, 32),
          ("up2.conv1", 32, 124, 2, False, False, 32), ("up3.conv0", 128, 252, 2, True, False, 32),
          ("up3.conv1", 32, 254, 2, False, True, 32)]
if len(sys.argv) > 3:   # experiments around up3.conv1: no trailing columns, no fused out conv, no padding
    LAYERS = [("u31", 32, 254, 2, False, True, 32), ("u31-notail", 32, 250, 2, False, True, 32), ("u31-nofuse", 32, 254, 2, False, False, 32),
              ("u31-nofuse-notail", 32, 250, 2, False, False, 32), ("pad0-fuse-252", 32, 254, 0, False, True, 32)]


def timed(fn):
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


tot_old = tot_new = 0.0
for name, ci, h, pad, derive, fuse, co in LAYERS:
    cs = ci // 4
    cin_t = 2 * cs if derive else ci
    x = torch.randn((n, cin_t // 8, h, h, 8), device=dev, generator=g).relu().to(torch.bfloat16)
    w9 = (torch.randn((9, ci, co), device=dev, generator=g) / (9 * ci) ** 0.5).to(torch.bfloat16).float()
    b = torch.randn(co, device=dev, generator=g) * 0.1
    ow, ob = torch.randn(32, device=dev, generator=g) * 0.3, torch.randn(1, device=dev, generator=g)
    ho = h + 2 * pad - 2
    wt, wr = packing.conv3x3_tc(w9), packing.conv3x3_tc_rows(w9)
    o0 = torch.empty((n, co // 8, ho, ho, 8), device=dev, dtype=torch.bfloat16)
    o1 = torch.empty_like(o0)
    img0, img1 = torch.empty((n, ho, ho), device=dev), torch.empty((n, ho, ho), device=dev)
    if derive:
        f_old = lambda: _lib.call("uncl_conv3x3_tc_skipcat", x, x.stride(0), wt, b, o0, o0.stride(0), _lib.BF16, n, cs, h, h, 32, pad, 1)
        f_new = lambda: _lib.call("uncl_conv3x3_tc_rows_skipcat", x, x.stride(0), wr, wt, b, o1, o1.stride(0), n, cs, h, h, pad, 1)
    elif fuse:
        f_old = lambda: _lib.call("uncl_conv3x3_tc", x, x.stride(0), wt, b, None, 0, _lib.BF16, n, ci, h, h, 32, pad, 1, 0, 1, ow, ob, img0, None)
        f_new = lambda: _lib.call("uncl_conv3x3_tc_rows", x, x.stride(0), wr, wt, b, None, 0, n, ci, h, h, 32, pad, 1, 0, 1, ow, ob, img1, None)
    else:
        f_old = lambda: _lib.call("uncl_conv3x3_tc", x, x.stride(0), wt, b, o0, o0.stride(0), _lib.BF16, n, ci, h, h, co, pad, 1, 0, 0, None, None, None, None)
        f_new = lambda: _lib.call("uncl_conv3x3_tc_rows", x, x.stride(0), wr, wt, b, o1, o1.stride(0), n, ci, h, h, co, pad, 1, 0, 0, None, None, None, None)
    t_old, t_new = timed(f_old), timed(f_new)
    a, r = (img1, img0) if fuse else (o1.float(), o0.float())
    err = ((a - r).norm() / r.norm()).item()
    gflop = 2 * 9 * ci * co * ho * ho * n / 1e9
    plan = packing.conv3x3_tc_rows_plan(n, ci, h, h, pad, derive, co=co)
    print("%-10s %4d tiles  old %8.1f us (%6.1f TFLOP/s)  rows %8.1f us (%6.1f TFLOP/s)  x%.2f  rel diff %.2e  plan %s"
          % (name, n, t_old, gflop / t_old * 1e3, t_new, gflop / t_new * 1e3, t_old / t_new, err, plan[:13]), flush=True)
    tot_old += t_old
    tot_new += t_new
print("sum: old %.1f us, rows %.1f us" % (tot_old, tot_new))

# inc.conv + inc.conv1: uncl_conv_first followed by the row kernel against the fused launch (front mode)
if len(sys.argv) <= 3:
    x = torch.rand((n, 1, 256, 256), device=dev, generator=g)
    w1 = torch.randn((32, 1, 3, 3), device=dev, generator=g) / 3
    b1 = torch.randn(32, device=dev, generator=g) * 0.1
    w9 = (torch.randn((9, 32, 32), device=dev, generator=g) / (9 * 32) ** 0.5).to(torch.bfloat16).float()
    b = torch.randn(32, device=dev, generator=g) * 0.1
    a0 = torch.empty((n, 4, 254, 254, 8), device=dev, dtype=torch.bfloat16)
    o0 = torch.empty((n, 4, 252, 252, 8), device=dev, dtype=torch.bfloat16)
    o1 = torch.empty_like(o0)
    wt, wr, cf, cfr = packing.conv3x3_tc(w9), packing.conv3x3_tc_rows(w9), packing.conv_first(w1), packing.conv_first_rows(w1, b1)

    def two():
        _lib.call("uncl_conv_first", x, cf, b1, a0, a0.stride(0), n, 256, 256, 32, 1, _lib.BF16)
        _lib.call("uncl_conv3x3_tc_rows", a0, a0.stride(0), wr, wt, b, o0, o0.stride(0), n, 32, 254, 254, 32, 0, 1, 0, 0, None, None, None, None)

    def one():
        _lib.call("uncl_conv_first_conv3x3_tc_rows", x, x.stride(0), cfr, wr, b, o1, o1.stride(0), n, 256, 256, 1, 0)

    t2, t1 = timed(two), timed(one)
    print("inc.conv + inc.conv1  %4d tiles  two launches %8.1f us  fused %8.1f us  x%.2f  rel diff %.2e"
          % (n, t2, t1, t2 / t1, ((o1.float() - o0.float()).norm() / o0.float().norm()).item()))
