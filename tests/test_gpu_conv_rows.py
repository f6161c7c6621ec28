"""Row kernel (conv_tc_rows.cu: ky taps merged into N, one image row per M block, rotating TMEM slots) against the
kernels it replaces on the C_out = 32 layers of the generator (unet_parts.py:57-87, :126-141, :183-193, :311-332)."""
import pytest
import torch

from uncltmo_b200 import _lib, packing

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _problem(ci, h, n, seed, co=32):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn((n, ci // 8, h, h, 8), device="cuda", generator=g).to(torch.bfloat16)
    w9 = (torch.randn((9, ci, co), device="cuda", generator=g) / (9 * ci) ** 0.5).to(torch.bfloat16).float()
    b = torch.randn(co, device="cuda", generator=g) * 0.1
    return g, x, w9, b


# (C_in, H = W, pad, images): the generator's own layers (inc.conv1 254 -> 252; up2.conv1 124 -> 126; up3.conv1 254 -> 256
# with 4 trailing columns; up3.conv0 over the materialised concat 252 -> 254 with 2 trailing columns), a width that leaves
# a 2-column tail after one band, tiny images (strips shorter than a row group), many images (several strips per CTA)
@pytest.mark.parametrize("ci,h,pad,n,co", [(32, 254, 0, 2, 32), (32, 124, 2, 2, 32), (32, 254, 2, 1, 32), (128, 252, 2, 1, 32),
                                           (32, 130, 0, 1, 32), (32, 20, 0, 3, 32), (64, 40, 2, 1, 32), (32, 7, 2, 5, 32),
                                           (32, 3, 0, 1, 32), (32, 60, 0, 300, 32), (96, 66, 2, 7, 32),
                                           # C_out = 64 (ring of eight 64-column groups): down0.conv0 126 -> 124, down0.conv1
                                           # 124 -> 122, a trailing-column case, tiny and many-image cases
                                           (32, 126, 0, 2, 64), (64, 124, 0, 2, 64), (64, 254, 2, 1, 64), (32, 9, 2, 4, 64),
                                           (64, 30, 0, 150, 64), (128, 33, 2, 3, 64)])
def test_row_kernel_matches_one_tap_and_fp32(ci, h, pad, n, co):
    g, x, w9, b = _problem(ci, h, n, ci * 1000 + h + pad + co, co)
    ho = h + 2 * pad - 2
    plan = packing.conv3x3_tc_rows_plan(n, ci, h, h, pad, co=co)
    assert plan[0] == 1 and plan[1] + plan[2] == ho
    ref = torch.empty((n, co // 8, ho, ho, 8), device="cuda", dtype=torch.float32)
    _lib.call("uncl_conv3x3_simt", x.float(), x.stride(0), w9, b, ref, ref.stride(0), n, ci, h, h, co, pad, 1, 0, _lib.F32)
    old = torch.empty((n, co // 8, ho, ho, 8), device="cuda", dtype=torch.bfloat16)
    wt = packing.conv3x3_tc(w9)
    _lib.call("uncl_conv3x3_tc", x, x.stride(0), wt, b, old, old.stride(0), _lib.BF16, n, ci, h, h, co, pad, 1, 0, 0,
              None, None, None, None)
    out = torch.full((n, co // 8, ho, ho, 8), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.call("uncl_conv3x3_tc_rows", x, x.stride(0), packing.conv3x3_tc_rows(w9), wt, b, out, out.stride(0), n, ci, h, h, co, pad,
              1, 0, 0, None, None, None, None)
    torch.cuda.synchronize()
    assert not torch.isnan(out.float()).any()
    # same bf16 operands, fp32 accumulation, one bf16 rounding of the result: only the summation order differs
    assert rel(out, ref) <= 4e-3 and rel(old, ref) <= 4e-3
    assert (out.float() - ref).abs().max().item() <= 1e-2 * max(1.0, ref.abs().max().item())
    assert rel(out, old) <= 2e-3


def test_row_kernel_skip_planes_64_channels():
    """down0.conv1 with materialised skip planes: o, o^2, sqrt(o + 1e-8) into a 256-channel concat buffer."""
    ci, co, h, n = 64, 64, 40, 3
    g, x, w9, b = _problem(ci, h, n, 4242, co)
    ho = h - 2
    wt, wr = packing.conv3x3_tc(w9), packing.conv3x3_tc_rows(w9)
    cat0 = torch.full((n, 32, ho, ho, 8), float("nan"), device="cuda", dtype=torch.bfloat16)
    cat1 = torch.full((n, 32, ho, ho, 8), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.call("uncl_conv3x3_tc", x, x.stride(0), wt, b, cat0, cat0.stride(0), _lib.BF16, n, ci, h, h, co, 0, 1, 1, 0,
              None, None, None, None)
    _lib.call("uncl_conv3x3_tc_rows", x, x.stride(0), wr, wt, b, cat1, cat1.stride(0), n, ci, h, h, co, 0, 1, 1, 0,
              None, None, None, None)
    torch.cuda.synchronize()
    for lo, hi in ((0, 8), (16, 24), (24, 32)):
        assert not torch.isnan(cat1[:, lo:hi].float()).any() and rel(cat1[:, lo:hi], cat0[:, lo:hi]) <= 2e-3
    assert torch.isnan(cat1[:, 8:16].float()).all()


@pytest.mark.parametrize("h,pad,n", [(254, 2, 2), (60, 0, 3)])
def test_row_kernel_fused_out_conv_and_skip_planes(h, pad, n):
    """The fused 1x1 out conv + sigmoid (all 32 channels of a pixel are in one thread) and the skip-plane emission
    (o, o^2, sqrt(o + 1e-8) into a 128-channel concat buffer) against the one-tap kernel's epilogues."""
    ci = 32
    g, x, w9, b = _problem(ci, h, n, 77 + h)
    ow, ob = torch.randn(32, device="cuda", generator=g) * 0.3, torch.randn(1, device="cuda", generator=g)
    ho = h + 2 * pad - 2
    wt, wr = packing.conv3x3_tc(w9), packing.conv3x3_tc_rows(w9)
    img0, logit0 = torch.empty((n, ho, ho), device="cuda"), torch.empty((n, ho, ho), device="cuda")
    _lib.call("uncl_conv3x3_tc", x, x.stride(0), wt, b, None, 0, _lib.BF16, n, ci, h, h, 32, pad, 1, 0, 1, ow, ob, img0, logit0)
    img = torch.full((n, ho, ho), float("nan"), device="cuda")
    logit = torch.full((n, ho, ho), float("nan"), device="cuda")
    _lib.call("uncl_conv3x3_tc_rows", x, x.stride(0), wr, wt, b, None, 0, n, ci, h, h, 32, pad, 1, 0, 1, ow, ob, img, logit)
    cat0 = torch.full((n, 16, ho, ho, 8), float("nan"), device="cuda", dtype=torch.bfloat16)
    cat1 = torch.full((n, 16, ho, ho, 8), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.call("uncl_conv3x3_tc", x, x.stride(0), wt, b, cat0, cat0.stride(0), _lib.BF16, n, ci, h, h, 32, pad, 1, 1, 0,
              None, None, None, None)
    _lib.call("uncl_conv3x3_tc_rows", x, x.stride(0), wr, wt, b, cat1, cat1.stride(0), n, ci, h, h, 32, pad, 1, 1, 0,
              None, None, None, None)
    torch.cuda.synchronize()
    assert not torch.isnan(logit).any() and not torch.isnan(img).any()
    assert (logit - logit0).abs().max().item() <= 1e-4 * max(1.0, logit0.abs().max().item())
    assert (img - img0).abs().max().item() <= 1e-5
    for lo, hi in ((0, 4), (8, 12), (12, 16)):
        assert not torch.isnan(cat1[:, lo:hi].float()).any()
        assert rel(cat1[:, lo:hi], cat0[:, lo:hi]) <= 2e-3
    assert torch.isnan(cat1[:, 4:8].float()).all()   # the up-sampled slice of the concat buffer stays untouched


@pytest.mark.parametrize("cs,h,pad,n", [(32, 252, 2, 1), (32, 40, 0, 2), (32, 5, 2, 3), (32, 122, 2, 9), (64, 122, 2, 2), (64, 9, 0, 4)])
def test_row_kernel_fused_skip_operators(cs, h, pad, n):
    """uncl_conv3x3_tc_rows_skipcat against uncl_conv3x3_tc_skipcat and against the conv over the materialised concat."""
    g = torch.Generator(device="cuda").manual_seed(h * 10 + pad)
    x = torch.randn((n, 2 * cs // 8, h, h, 8), device="cuda", generator=g)
    x[:, :cs // 8] = x[:, :cs // 8].relu()
    x = x.to(torch.bfloat16)
    skip = x[:, :cs // 8].float()
    full = torch.cat([x, (skip * skip).to(torch.bfloat16), torch.sqrt(skip + 1e-8).to(torch.bfloat16)], dim=1).contiguous()
    w9 = (torch.randn((9, 4 * cs, 32), device="cuda", generator=g) / (9 * 4 * cs) ** 0.5).to(torch.bfloat16).float()
    b = torch.randn(32, device="cuda", generator=g) * 0.1
    ho = h + 2 * pad - 2
    assert packing.conv3x3_tc_rows_plan(n, 4 * cs, h, h, pad, derive=True)[0] == 1
    wt, wr = packing.conv3x3_tc(w9), packing.conv3x3_tc_rows(w9)
    ref = torch.empty((n, 4, ho, ho, 8), device="cuda", dtype=torch.float32)
    _lib.call("uncl_conv3x3_simt", full.float(), full.stride(0), w9, b, ref, ref.stride(0), n, 4 * cs, h, h, 32, pad, 1, 0, _lib.F32)
    old = torch.empty((n, 4, ho, ho, 8), device="cuda", dtype=torch.bfloat16)
    _lib.call("uncl_conv3x3_tc_skipcat", x, x.stride(0), wt, b, old, old.stride(0), _lib.BF16, n, cs, h, h, 32, pad, 1)
    out = torch.full((n, 4, ho, ho, 8), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.call("uncl_conv3x3_tc_rows_skipcat", x, x.stride(0), wr, wt, b, out, out.stride(0), n, cs, h, h, pad, 1)
    torch.cuda.synchronize()
    assert not torch.isnan(out.float()).any()
    assert rel(out, ref) <= 4e-3 and rel(out, old) <= 2e-3


def test_row_kernel_network_matches_older_kernels():
    """The bf16 generator with the row kernel on inc.conv1 / up2.conv1 / up3.conv* (default) against the same network on
    the older kernels, with fused and with materialised skip planes."""
    from uncltmo_b200.generator import UNet
    from uncltmo_b200.weights import make_generator_state_dict
    g_args = (1, 1, "sigmoid", 4, 4, "square_and_square_root", 32, 0, "unet", 0, 0, "none", "none", "relu", 1, "replicate", 2)
    net = UNet(*g_args, up_mode=0, precision="bf16").cuda().eval()
    net.load_state_dict(make_generator_state_dict())
    x = torch.rand((3, 1, 256, 256), device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    outs = {}
    with torch.no_grad():
        for fused in (True, False):
            for rows in (True, False):
                net.fused_skip, net.row_kernel = fused, rows
                outs[(fused, rows)] = net.tonemap_tiles(x).clone()
    torch.cuda.synchronize()
    assert all(k + "_rows" in net.packed() for k in ("inc1", "d0_0", "d0_1", "u2_0", "u2_1", "u3_0", "u3_1"))
    for fused in (True, False):
        assert rel(outs[(fused, True)], outs[(fused, False)]) <= 2e-3


@pytest.mark.parametrize("ci,co,h,pad", [(32, 32, 70, 0), (64, 64, 70, 0), (128, 32, 70, 2), (32, 64, 70, 2)])
def test_row_kernel_is_batch_invariant(ci, co, h, pad):
    """Every output bit is independent of how many images share the launch (the strip height follows the batch size, the
    summation order must not): image 0 alone, in a batch of 7 and in a batch of 300."""
    g, x, w9, b = _problem(ci, h, 300, 99 + ci + co, co)
    ho = h + 2 * pad - 2
    wt, wr = packing.conv3x3_tc(w9), packing.conv3x3_tc_rows(w9)
    outs = []
    for n in (1, 7, 300):
        out = torch.empty((n, co // 8, ho, ho, 8), device="cuda", dtype=torch.bfloat16)
        _lib.call("uncl_conv3x3_tc_rows", x[:n], x.stride(0), wr, wt, b, out, out.stride(0), n, ci, h, h, co, pad, 1, 0, 0,
                  None, None, None, None)
        outs.append(out[0].clone())
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


@pytest.mark.parametrize("n,h0,emit", [(2, 256, 0), (1, 256, 1), (3, 60, 0), (150, 40, 0), (1, 130, 1)])
def test_row_kernel_front_mode_fuses_the_first_conv(n, h0, emit):
    """uncl_conv_first_conv3x3_tc_rows (inc.conv computed into the stage ring of inc.conv1's launch, three-term bf16 split)
    against uncl_conv_first (fp32 FMA, bf16 output) followed by the row kernel, and against the fp32 convs."""
    g = torch.Generator(device="cuda").manual_seed(11 * n + h0)
    x = torch.rand((n, 1, h0, h0), device="cuda", generator=g)
    w1 = torch.randn((32, 1, 3, 3), device="cuda", generator=g) / 3
    b1 = torch.randn(32, device="cuda", generator=g) * 0.1
    w9 = (torch.randn((9, 32, 32), device="cuda", generator=g) / (9 * 32) ** 0.5).to(torch.bfloat16).float()
    b = torch.randn(32, device="cuda", generator=g) * 0.1
    ha, ho = h0 - 2, h0 - 4
    a0 = torch.empty((n, 4, ha, ha, 8), device="cuda", dtype=torch.bfloat16)
    _lib.call("uncl_conv_first", x, packing.conv_first(w1), b1, a0, a0.stride(0), n, h0, h0, 32, 1, _lib.BF16)
    cb = 16 if emit else 4
    two = torch.full((n, cb, ho, ho, 8), float("nan"), device="cuda", dtype=torch.bfloat16)
    one = torch.full((n, cb, ho, ho, 8), float("nan"), device="cuda", dtype=torch.bfloat16)
    wt, wr = packing.conv3x3_tc(w9), packing.conv3x3_tc_rows(w9)
    _lib.call("uncl_conv3x3_tc_rows", a0, a0.stride(0), wr, wt, b, two, two.stride(0), n, 32, ha, ha, 32, 0, 1, emit, 0,
              None, None, None, None)
    _lib.call("uncl_conv_first_conv3x3_tc_rows", x, x.stride(0), packing.conv_first_rows(w1, b1), wr, b, one, one.stride(0), n,
              h0, h0, 1, emit)
    ref0 = torch.relu(torch.nn.functional.conv2d(x, w1, b1))
    w2 = w9.reshape(3, 3, 32, 32).permute(3, 2, 0, 1).contiguous()
    ref = torch.relu(torch.nn.functional.conv2d(ref0.to(torch.bfloat16).float(), w2, b))       # [n, 32, ho, ho]
    ref_blocked = ref.reshape(n, 4, 8, ho, ho).permute(0, 1, 3, 4, 2)
    torch.cuda.synchronize()
    planes = ((0, 4), (8, 12), (12, 16)) if emit else ((0, 4),)
    for lo, hi in planes:
        assert not torch.isnan(one[:, lo:hi].float()).any()
        assert rel(one[:, lo:hi], two[:, lo:hi]) <= 3e-3
    if emit:
        assert torch.isnan(one[:, 4:8].float()).all()
    assert rel(one[:, :4], ref_blocked) <= 5e-3 and rel(two[:, :4], ref_blocked) <= 5e-3


def test_front_mode_network_matches_unfused():
    from uncltmo_b200.generator import UNet
    from uncltmo_b200.weights import make_generator_state_dict
    g_args = (1, 1, "sigmoid", 4, 4, "square_and_square_root", 32, 0, "unet", 0, 0, "none", "none", "relu", 1, "replicate", 2)
    net = UNet(*g_args, up_mode=0, precision="bf16").cuda().eval()
    net.load_state_dict(make_generator_state_dict())
    x = torch.rand((5, 1, 256, 256), device="cuda", generator=torch.Generator(device="cuda").manual_seed(6))
    with torch.no_grad():
        net.fused_first = True
        a = net.tonemap_tiles(x).clone()
        a1 = net.tonemap_tiles(x[:1]).clone()
        net.fused_first = False
        b = net.tonemap_tiles(x).clone()
    torch.cuda.synchronize()
    assert rel(a, b) <= 2e-3 and torch.equal(a[:1], a1)
