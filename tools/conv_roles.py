"""Diagnostic: where each role of the tensor-core conv kernel spends its cycles, per layer of one 1080p frame (60 tiles).
Uses uncl_conv_tc_set_debug; prints per launch the share of cycles each warp role spent waiting."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from uncltmo_b200 import _lib, generator as G
from uncltmo_b200.weights import make_generator_state_dict

G_ARGS = (1, 1, "sigmoid", 4, 4, "square_and_square_root", 32, 0, "unet", 0, 0, "none", "none", "relu", 1, "replicate", 2)
net = G.UNet(*G_ARGS, up_mode=0, precision="bf16").cuda().eval()
net.load_state_dict(make_generator_state_dict())
x = torch.rand(60, 1, 256, 256, device="cuda")
cnt = torch.zeros(12, dtype=torch.int64, device="cuda")
rows = []
orig = G.call


def traced(name, *args):
    if name in ("uncl_conv3x3_tc", "uncl_convT2x2_tc", "uncl_pw_conv_tc"):
        cnt.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig(name, *args)
        e1.record()
        torch.cuda.synchronize()
        c = cnt.tolist()
        ints = [a for a in args if isinstance(a, int)]
        rows.append((name, ints[:9], e0.elapsed_time(e1) * 1e3, c))
        return r
    return orig(name, *args)


with torch.no_grad():
    net.tonemap_tiles(x)
    torch.cuda.synchronize()
    _lib.lib().uncl_conv_tc_set_debug(cnt.data_ptr())
    G.call = traced
    net.tonemap_tiles(x)
    G.call = orig
    _lib.lib().uncl_conv_tc_set_debug(None)
print("%-18s %-44s %8s | producer: wait-empty | mma: wait-full wait-acc busy | epi: wait-acc" % ("kernel", "int args", "us"))
for name, ints, us, c in rows:
    prod, p_we, mma, m_wf, m_wa, epi, e_wa, n, ns = c[:9]
    print("%-18s %-44s %8.1f | %5.1f%% | %5.1f%% %5.1f%% %5.1f%% | %5.1f%%  (ctas %d, mma cycles/cta %.0f, SM clock %.0f MHz)" % (
        name, str(ints), us, 100 * p_we / max(prod, 1), 100 * m_wf / max(mma, 1), 100 * m_wa / max(mma, 1),
        100 * (mma - m_wf - m_wa) / max(mma, 1), 100 * e_wa / max(epi, 1), n, mma / max(n, 1), 1e3 * mma / max(ns, 1)))
