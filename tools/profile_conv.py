"""Runs a few tcgen05 conv launches at the generator's 60-tile shapes (for `ncu --set full -k regex:conv3x3_tc`)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uncltmo_b200 import _lib, packing  # noqa: E402

CASES = {  # name: (C_in, C_out, H, pad, emit_skip, fuse_outc) - the 17 tensor-core 3x3 layers of the generator
    "inc1": (32, 32, 254, 0, 1, 0), "d0_0": (32, 64, 126, 0, 0, 0), "d0_1": (64, 64, 124, 0, 1, 0),
    "d1_0": (64, 128, 61, 0, 0, 0), "d1_1": (128, 128, 59, 0, 1, 0), "d2_0": (128, 256, 28, 0, 0, 0),
    "d2_1": (256, 256, 26, 0, 1, 0), "d3_0": (256, 256, 12, 0, 0, 0), "bott": (256, 256, 10, 2, 0, 0),
    "u0_0": (1024, 128, 24, 2, 0, 0), "u0_1": (128, 128, 26, 2, 0, 0), "u1_0": (512, 64, 57, 2, 0, 0),
    "u1_1": (64, 64, 59, 2, 0, 0), "u2_0": (256, 32, 122, 2, 0, 0), "u2_1": (32, 32, 124, 2, 0, 0),
    "u3_0": (128, 32, 252, 2, 0, 0), "u3_1": (32, 32, 254, 2, 0, 1),
    # fused skip operators (uncl_conv3x3_tc_skipcat): emit = 2 marks the consumer, inc1n / d0_1n the producers without skip planes
    "u3_0f": (128, 32, 252, 2, 2, 0), "u2_0f": (256, 32, 122, 2, 2, 0), "inc1n": (32, 32, 254, 0, 0, 0), "d0_1n": (64, 64, 124, 0, 0, 0),
}


def run(name, n=60, reps=1):
    ci, co, h, pad, emit, fuse = CASES[name]
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((n, ci // 8, h, h, 8), device="cuda", generator=g).to(torch.bfloat16)
    w9 = torch.randn((9, ci, co), device="cuda", generator=g) / (9 * ci) ** 0.5
    b = torch.zeros(co, device="cuda")
    ho = h + 2 * pad - 2
    out = torch.empty((n, (4 if emit else 1) * co // 8, ho, ho, 8), device="cuda", dtype=torch.bfloat16)
    wp = packing.conv3x3_tc(w9)
    ow, ob = torch.randn(co, device="cuda"), torch.zeros(1, device="cuda")
    img = torch.empty((n, ho, ho), device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cnt = None
    if os.environ.get("PROFILE_DBG"):   # probes library only: per-role cycle counters + the SM clock the launch ran at
        cnt = torch.zeros(12, dtype=torch.int64, device="cuda")
        _lib.lib().uncl_conv_tc_set_debug(cnt.data_ptr())
    skipcat = emit == 2
    if skipcat:
        emit = 0
        x = x[:, :ci // 16].contiguous()
    for i in range(reps + 1):
        if i == 1:
            e0.record()
        if skipcat:
            _lib.call("uncl_conv3x3_tc_skipcat", x, x.stride(0), wp, b, out, out.stride(0), _lib.BF16, n, ci // 4, h, h, co, pad, 1)
            continue
        _lib.call("uncl_conv3x3_tc", x, x.stride(0), wp, b, None if fuse else out, out.stride(0), _lib.BF16, n, ci, h, h, co, pad, 1,
                  emit, fuse, ow if fuse else None, ob if fuse else None, img if fuse else None, None)
    e1.record()
    torch.cuda.synchronize()
    if reps:
        ms = e0.elapsed_time(e1) / reps
        fl = 2.0 * 9 * ci * co * ho * ho * n
        extra = ""
        if cnt is not None:
            c = cnt.tolist()
            _lib.lib().uncl_conv_tc_set_debug(None)
            extra = "  mma cycles/cta/launch %.0f  SM clock %.0f MHz  mma wait-full %.1f%% wait-acc %.1f%%" % (
                c[2] / max(c[7], 1), 1e3 * c[2] / max(c[8], 1), 100 * c[3] / max(c[2], 1), 100 * c[4] / max(c[2], 1))
            if c[9]:
                extra += "  skip-op warps: wait-stage %.1f%% wait-skip %.1f%% busy %.1f%%" % (
                    100 * c[10] / c[9], 100 * c[11] / c[9], 100 * (c[9] - c[10] - c[11]) / c[9])
        print("%-5s %8.1f us  %7.1f TFLOP/s%s" % (name, ms * 1e3, fl / ms / 1e9, extra), flush=True)


if __name__ == "__main__":
    names = sys.argv[1].split(",") if len(sys.argv) > 1 and sys.argv[1] != "all" else list(CASES)
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    for nm in names:
        run(nm, reps=reps)
