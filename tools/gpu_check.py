"""Diagnostic run on a GPU box: per-layer errors of the CUDA generator (fp32 and bf16) against the CPU oracle."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import golden_inputs as gi  # noqa: E402
import oracle  # noqa: E402
from uncltmo_b200 import _lib  # noqa: E402
from uncltmo_b200.generator import UNet, blocked_to_nchw  # noqa: E402
from uncltmo_b200.weights import make_generator_state_dict  # noqa: E402

torch.set_grad_enabled(False)


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def conv_unit_tests():
    """tcgen05 conv vs the CUDA-core conv on identical bf16 inputs, one case per layer geometry."""
    from uncltmo_b200 import packing
    cases = [(32, 32, 254, 0, 2), (32, 64, 126, 0, 2), (64, 64, 124, 0, 1), (64, 128, 61, 0, 2), (128, 128, 59, 0, 1),
             (128, 256, 28, 0, 2), (256, 256, 26, 0, 1), (256, 256, 12, 0, 3), (256, 256, 10, 2, 3),
             (1024, 128, 24, 2, 1), (128, 128, 26, 2, 1), (512, 64, 57, 2, 1), (64, 64, 59, 2, 2),
             (256, 32, 122, 2, 1), (32, 32, 124, 2, 2), (128, 32, 252, 2, 1), (32, 32, 254, 2, 1)]
    g = torch.Generator(device="cuda").manual_seed(0)
    for ci, co, h, pad, n in cases:
        x = torch.randn((n, ci // 8, h, h, 8), device="cuda", generator=g).to(torch.bfloat16)
        w9 = torch.randn((9, ci, co), device="cuda", generator=g) * (1.0 / (9 * ci) ** 0.5)
        w9 = w9.to(torch.bfloat16).float()
        b = torch.randn(co, device="cuda", generator=g) * 0.1
        ho = h + 2 * pad - 2
        ref = torch.empty((n, co // 8, ho, ho, 8), device="cuda", dtype=torch.bfloat16)
        out = torch.full((n, co // 8, ho, ho, 8), float("nan"), device="cuda", dtype=torch.bfloat16)
        _lib.call("uncl_conv3x3_simt", x, x.stride(0), w9, b, ref, ref.stride(0), n, ci, h, h, co, pad, 1, 0, _lib.BF16)
        _lib.call("uncl_conv3x3_tc", x, x.stride(0), packing.conv3x3_tc(w9), b, out, out.stride(0), _lib.BF16, n, ci, h, h, co, pad,
                  1, 0, 0, None, None, None, None)
        torch.cuda.synchronize()
        nan = torch.isnan(out.float()).sum().item()
        print("conv_tc ci=%4d co=%3d h=%3d pad=%d n=%d  rel=%.3e  maxabs=%.3e  nan=%d" %
              (ci, co, h, pad, n, rel(out.float(), ref.float()), (out.float() - ref.float()).abs().max().item(), nan), flush=True)


def main():
    l = _lib.lib()
    print("lib", l.uncl_arch().decode(), l.uncl_version(), torch.cuda.get_device_name(0))
    d = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.call("uncl_probe_device", d)
    print("probe", d.item())
    sd = make_generator_state_dict()
    x = gi.generator_input()
    t0 = time.time()
    o_out, o_up, inter = oracle.unet_forward(sd, x, return_all=True)
    print("oracle forward %.2fs" % (time.time() - t0))
    for prec in ("fp32", "bf16"):
        if prec == "bf16":
            try:
                conv_unit_tests()
            except Exception as e:  # noqa: BLE001
                print("conv unit tests FAILED:", e)
        net = UNet(1, 1, "sigmoid", 4, 4, "square_and_square_root", 32, 0, "unet", 0, 0, "none", "none", "relu", 1,
                   "replicate", 2, up_mode=0, precision=prec).cuda().eval()
        net.load_state_dict(sd)
        keep = {}
        out, up, logit, _ = net._run_frame(x.cuda(), want_logit=True, keep=keep)
        torch.cuda.synchronize()
        print("[%s] out rel=%.3e logit rel=%.3e up_x rel=%.3e" %
              (prec, rel(out, o_out), rel(logit, inter["logit"]), rel(blocked_to_nchw(up), o_up)))
        for i in range(4):
            c = inter["skips"][i].shape[1]
            sk = blocked_to_nchw(keep["skips"][i][:, :c // 8].contiguous())
            print("   skip%d rel=%.3e" % (i, rel(sk, inter["skips"][i])))
        print("   x4 rel=%.3e  gcn rel=%.3e" % (rel(blocked_to_nchw(keep["x4"]), inter["skips"][4]),
                                               rel(blocked_to_nchw(keep["gcn"]), inter["gcn"])))
        for i in range(4):
            print("   up%d rel=%.3e" % (i, rel(blocked_to_nchw(keep["ups"][i]), inter["ups"][i])))
        # KNN agreement
        _, oidx = oracle.gcn_block(sd, inter["skips"][4], return_idx=True)
        mine = keep["idx"].cpu().long().sort(dim=-1)[0]
        print("   knn index agreement: %.4f" % (mine == oidx.sort(dim=-1)[0]).float().mean().item())
        if prec == "bf16":
            out2 = net.tonemap_tiles(x.cuda())
            torch.cuda.synchronize()
            print("   fused-outc path rel=%.3e" % rel(out2, o_out))
        # timing at 60 tiles
        xb = torch.rand(60, 1, 256, 256, device="cuda")
        for _ in range(2):
            net.tonemap_tiles(xb)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            net.tonemap_tiles(xb)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print("[%s] 60 tiles: %.2f ms  -> %.1f TFLOP/s" % (prec, ms, 60 * 18.2858 / ms))


if __name__ == "__main__":
    main()
