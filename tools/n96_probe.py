"""Probe: cycles per tcgen05.mma (SS mode) of the conv kernel as a function of N (C_out), large image, K = 128."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from uncltmo_b200 import _lib, packing
cnt = torch.zeros(8, dtype=torch.int64, device="cuda")
n, ci, h = 60, 128, 252
x = torch.randn((n, ci // 8, h, h, 8), device="cuda").bfloat16()
for co in (32, 64, 96, 128) + ((256,) if os.environ.get('UNCL_PROBE_NT256') else ()):
    w9 = torch.randn((9, ci, co), device="cuda") * 0.03
    wp = packing.conv3x3_tc(w9) if co != 256 else torch.randn((1, ci // 16, 9, 2, 256, 8), device='cuda').bfloat16() * 0.03
    b = torch.zeros(co, device="cuda")
    out = torch.empty((n, co // 8, h + 2, h + 2, 8), device="cuda", dtype=torch.bfloat16)
    for it in range(2):
        cnt.zero_()
        _lib.lib().uncl_conv_tc_set_debug(cnt.data_ptr() if it else None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call("uncl_conv3x3_tc", x, x.stride(0), wp, b, out, out.stride(0), _lib.BF16, n, ci, h, h, co, 2, _lib.ACT_RELU, 0, 0,
                  None, None, None, None)
        e1.record()
        torch.cuda.synchronize()
    _lib.lib().uncl_conv_tc_set_debug(None)
    c = cnt.tolist()
    NT = min(co, 128) if co != 256 else 256
    PW = 86  # 254 / 3 bands + 2
    mmas_total = n * 254 * 254 / 128 * 1.05 * (ci // 16) * 9
    print("C_out %3d: %.1f us, mma-warp cycles/cta %.0f, wait-full %.1f%% wait-acc %.1f%%, ~%.1f cycles per MMA (floor %d)" % (
        co, e0.elapsed_time(e1) * 1e3, c[2] / c[7], 100 * c[3] / c[2], 100 * c[4] / c[2], c[2] / c[7] / (mmas_total / 148), NT // 2))
