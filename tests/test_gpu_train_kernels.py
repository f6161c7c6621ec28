"""Unit parity of the bf16 training-path kernels (train_bf16.cu, the fused-mask data gradient, the tensor-core weight
gradients) through the C ABI against float64 torch autograd on the CPU.  Operands are pre-rounded to bf16 so that only
the accumulation order (fp32 vs float64) and the final bf16 store differ."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle
from uncltmo_b200 import _lib, packing
from uncltmo_b200._lib import BF16, F32, call

pytestmark = pytest.mark.gpu


def to_blocked(x, dtype=torch.bfloat16):  # NCHW -> [N, C/8, H, W, 8]
    n, c, h, w = x.shape
    return x.reshape(n, c // 8, 8, h, w).permute(0, 1, 3, 4, 2).contiguous().to(dtype).cuda()


def from_blocked(x):
    n, cb, h, w, _ = x.shape
    return x.float().permute(0, 1, 4, 2, 3).reshape(n, cb * 8, h, w).cpu()


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def rnd(*shape, seed=0, scale=1.0):
    t = torch.from_numpy((np.random.default_rng(seed).standard_normal(shape) * scale).astype(np.float32))
    return t.to(torch.bfloat16).float()    # bf16-representable


@pytest.mark.parametrize("ci,co,h,w,pad", [(32, 32, 70, 73, 0), (32, 32, 21, 19, 2), (128, 32, 66, 67, 2), (256, 32, 30, 33, 2),
                                          (64, 32, 17, 130, 0), (32, 32, 254, 254, 0), (32, 64, 37, 40, 0),
                                          (128, 128, 29, 31, 2), (64, 64, 60, 61, 2)])
def test_conv3x3_weight_gradient_tensor_cores(ci, co, h, w, pad):
    """uncl_conv3x3_wgrad_tc: the transposed C_out = 32 kernel (filter rows stacked in M from a [row][block][pixel] dZ
    tile) and the general kernel, ragged extents, valid and full (ConvTranspose) padding."""
    n = 2
    x, dz = rnd(n, ci, h, w, seed=1), rnd(n, co, h + 2 * pad - 2, w + 2 * pad - 2, seed=2)
    wt = torch.zeros(co, ci, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double(), wt, padding=pad).backward(dz.double())
    want = wt.grad.permute(2, 3, 1, 0).reshape(9, ci, co)          # [tap][ci][co]
    xb, zb = to_blocked(x), to_blocked(dz)
    dw = torch.zeros(9, ci, co, device="cuda")
    call("uncl_conv3x3_wgrad_tc", xb, xb.stride(0), zb, dw, n, ci, h, w, co, pad)
    assert rel(dw, want) < 2e-6
    call("uncl_conv3x3_wgrad_tc", xb, xb.stride(0), zb, dw, n, ci, h, w, co, pad)      # accumulates
    assert rel(dw, 2 * want) < 2e-6


@pytest.mark.parametrize("ci,co,h,w", [(32, 128, 126, 126), (64, 256, 61, 61), (256, 1024, 12, 12), (128, 512, 28, 28)])
def test_pointwise_weight_gradient_tensor_cores(ci, co, h, w):
    n = 2
    x, dz = rnd(n, ci, h, w, seed=3), rnd(n, co, h, w, seed=4)
    want = torch.einsum("nihw,nohw->io", x.double(), dz.double())
    xb, zb = to_blocked(x), to_blocked(dz)
    dw = torch.zeros(ci, co, device="cuda")
    call("uncl_pw_wgrad_tc", xb, xb.stride(0), zb, zb.stride(0), dw, n, ci, co, h, w)
    assert rel(dw, want) < 2e-6


@pytest.mark.parametrize("ci,co,h,pad,f32out", [(32, 32, 40, 0, False), (32, 32, 40, 2, True), (64, 64, 30, 2, False),
                                               (32, 128, 37, 0, False), (128, 128, 20, 0, False), (256, 256, 12, 0, False)])
def test_data_gradient_with_fused_relu_mask(ci, co, h, pad, f32out):
    """uncl_conv3x3_tc_dgrad: dX = corr(dZ, transposed taps, pad 2 - p) * (mask > 0), both tensor-core kernels (one tap per
    MMA and kx-merged), bf16 and fp32 output.  (ci, co) are those of the DATA-GRADIENT GEMM: dZ has ci channels."""
    n = 2
    fwd_pad = 2 - pad
    w = rnd(ci, co, 3, 3, seed=5, scale=(9 * ci) ** -0.5)      # forward conv: co -> ci channels, weight [ci][co][3][3]
    dz = rnd(n, ci, h, h + 3, seed=6)
    xin = torch.zeros(n, co, h + 2 * pad - 2, h + 3 + 2 * pad - 2, dtype=torch.float64, requires_grad=True)
    F.conv2d(xin, w.double(), padding=fwd_pad).backward(dz.double())
    mask = rnd(*xin.shape, seed=7).clamp(min=0)                  # a post-ReLU activation: zeros and positives
    want = xin.grad * (mask.double() > 0)
    w9 = packing.conv3x3_taps(w, False)                          # forward taps [9][co][ci]
    wd = packing.conv3x3_tc(packing.conv3x3_dgrad_taps_layout(w9)).cuda()
    zb, mb = to_blocked(dz), to_blocked(mask)
    out = torch.full((n, co // 8, want.shape[2], want.shape[3], 8), float("nan"), device="cuda",
                     dtype=torch.float32 if f32out else torch.bfloat16)
    call("uncl_conv3x3_tc_dgrad", zb, zb.stride(0), wd, mb, mb.stride(0), out, out.stride(0), F32 if f32out else BF16, n, ci,
         h, h + 3, co, pad)
    got = from_blocked(out)
    assert not torch.isnan(got).any()
    assert rel(got, want) < (2e-6 if f32out else 3e-3)
    assert (got[mask == 0] == 0).all()


def test_skip_pool_backward():
    """uncl_skip_pool_bwd against autograd of relu -> {cat([x2, up, x2^2, sqrt(x2+1e-8)]), max_pool2d}."""
    n, c, h, w = 2, 16, 23, 26
    z = rnd(n, c, h, w, seed=8)
    gcat = rnd(n, 4 * c, h, w, seed=9)
    gpool = rnd(n, c, h // 2, w // 2, seed=10)
    zr = z.double().requires_grad_(True)
    x2 = F.relu(zr)
    cat = torch.cat([x2, torch.zeros_like(x2), x2 * x2, torch.pow(x2 + 1e-8, 0.5)], 1)
    ((cat * gcat.double()).sum() + (F.max_pool2d(x2, 2) * gpool.double()).sum()).backward()
    catb = torch.zeros((n, 4 * c // 8, h, w, 8), device="cuda", dtype=torch.bfloat16)
    catb[:, :c // 8] = to_blocked(F.relu(z))
    dz = torch.empty((n, c // 8, h, w, 8), device="cuda", dtype=torch.bfloat16)
    db = torch.zeros(c, device="cuda")
    call("uncl_skip_pool_bwd", catb, catb.stride(0), to_blocked(gcat), to_blocked(gpool), dz, db, n, c, h, w)
    got = from_blocked(dz)
    # elements with tiny positive x2 see 0.5 / sqrt(x2 + 1e-8) up to 5000: compare relative to the local magnitude
    assert rel(got, zr.grad) < 4e-3
    assert (got[z <= 0] == 0).all()
    assert rel(db, got.double().sum(dim=(0, 2, 3))) < 1e-5
    # without a concat gradient (or without a pool gradient) the other path alone
    call("uncl_skip_pool_bwd", catb, catb.stride(0), None, to_blocked(gpool), dz, None, n, c, h, w)
    zr.grad = None
    (F.max_pool2d(F.relu(zr), 2) * gpool.double()).sum().backward()
    assert rel(from_blocked(dz), zr.grad) < 4e-3


def test_skip_pool_backward_with_recurrence():
    """uncl_skip_pool_bwd_rec: the video generator pools cat(prev[:, :r], x2[:, r:]) (Unet.py:244).  Against autograd:
    gradient to x2's producer (with the gradient the next frame sends to the own slice) and to prev's first r channels."""
    n, c, h, w, r = 2, 64, 22, 25, 2
    z, zp = rnd(n, c, h, w, seed=30), rnd(n, c, h, w, seed=31)
    gcat, gpool = rnd(n, 4 * c, h, w, seed=32), rnd(n, c, h // 2, w // 2, seed=33)
    gstate = rnd(n, r, h, w, seed=34)
    zr, pr = z.double().requires_grad_(True), F.relu(zp.double()).requires_grad_(True)
    x2 = F.relu(zr)
    cat = torch.cat([x2, torch.zeros_like(x2), x2 * x2, torch.pow(x2 + 1e-8, 0.5)], 1)
    fea = torch.cat([pr[:, :r], x2[:, r:]], 1)
    ((cat * gcat.double()).sum() + (F.max_pool2d(fea, 2) * gpool.double()).sum() + (x2[:, :r] * gstate.double()).sum()).backward()
    catb = torch.zeros((n, 4 * c // 8, h, w, 8), device="cuda", dtype=torch.bfloat16)
    catb[:, :c // 8] = to_blocked(F.relu(z))
    prevb = to_blocked(F.relu(zp))
    ds = torch.zeros(n, 8, h, w)
    ds[:, :r] = gstate
    dz = torch.empty((n, c // 8, h, w, 8), device="cuda", dtype=torch.bfloat16)
    dprev = torch.full((n, 1, h, w, 8), float("nan"), device="cuda", dtype=torch.bfloat16)
    db = torch.zeros(c, device="cuda")
    call("uncl_skip_pool_bwd_rec", catb, catb.stride(0), to_blocked(gcat), to_blocked(gpool), dz, db, n, c, h, w, prevb,
         prevb.stride(0), r, dprev, to_blocked(ds))
    got, gotp = from_blocked(dz), from_blocked(dprev)
    assert rel(got, zr.grad) < 4e-3
    assert rel(gotp[:, :r], pr.grad[:, :r]) < 4e-3 and (gotp[:, r:] == 0).all()


def test_decoder_splice_gradient_fixup():
    """uncl_splice_grad: the up-convolution read cat(prev[:, :r], up[:, r:]) (Unet.py:270)."""
    n, c, h, r = 2, 64, 15, 2
    g = rnd(n, c, h, h, seed=35)
    own = rnd(n, c, h, h, seed=36).clamp(min=0)
    gstate = rnd(n, 8, h, h, seed=37)
    dz = to_blocked(g)
    dprev = torch.full((n, 1, h, h, 8), float("nan"), device="cuda", dtype=torch.bfloat16)
    ownb = to_blocked(own)
    call("uncl_splice_grad", dz, dz.stride(0), ownb, ownb.stride(0), r, dprev, to_blocked(gstate), n, h * h)
    want = g.clone()
    want[:, :r] = (own[:, :r] > 0) * gstate[:, :r]
    assert rel(from_blocked(dz), want) < 4e-3
    p = from_blocked(dprev)
    assert torch.equal(p[:, :r], g[:, :r]) and (p[:, r:] == 0).all()
    # no splice at this tensor (first frame): only the next frame's gradient is added
    dz = to_blocked(g)
    call("uncl_splice_grad", dz, dz.stride(0), ownb, ownb.stride(0), r, None, to_blocked(gstate), n, h * h)
    want = g.clone()
    want[:, :r] += (own[:, :r] > 0) * gstate[:, :r]
    assert rel(from_blocked(dz), want) < 4e-3


@pytest.mark.parametrize("c,h,h2", [(32, 13, 26), (64, 28, 57)])
def test_upconv_space_to_depth_backward(c, h, h2):
    """uncl_convT2x2_s2d_bf16: the k2 s2 up-convolution's output gradient (a channel slice of the concat gradient, with the
    replicate pad of unet_parts.py:283-299 folded back) as a [N, 4C, H, W] tensor + the bias gradient."""
    n = 2
    dcat = rnd(n, 4 * c, h2, h2, seed=11)
    x = torch.zeros(n, c, h, h, dtype=torch.float64)
    wt = rnd(c, c, 2, 2, seed=12).double().requires_grad_(True)
    bias = torch.zeros(c, dtype=torch.float64, requires_grad=True)
    xin = rnd(n, c, h, h, seed=13).double()
    y = F.conv_transpose2d(xin, wt, bias, stride=2)
    d = h2 - 2 * h
    if d:
        y = F.pad(y, (d // 2, d - d // 2, d // 2, d - d // 2), mode="replicate")
    y.backward(dcat[:, c:2 * c].double())
    db_b = to_blocked(dcat)
    s2d = torch.empty((n, 4 * c // 8, h, h, 8), device="cuda", dtype=torch.bfloat16)
    db = torch.zeros(c, device="cuda")
    call("uncl_convT2x2_s2d_bf16", db_b[:, c // 8:], db_b.stride(0), s2d, db, n, c, h, h, h2, h2)
    s = from_blocked(s2d).double()                                  # [n, pos*C + co, h, w]
    want_w = torch.einsum("nihw,nphw->ip", xin, s).reshape(c, 2, 2, c).permute(0, 3, 1, 2)
    assert rel(want_w, wt.grad) < 5e-3 and rel(db, bias.grad) < 5e-3
    del x


def test_out_conv_feature_backward():
    """uncl_outc_feat_bwd against autograd of sigmoid(conv1x1(relu(z))) + a feature-path gradient on relu(z)."""
    n, c, h = 2, 32, 24
    z = rnd(n, c, h, h, seed=14)
    w, b = rnd(1, c, 1, 1, seed=15, scale=0.2), torch.tensor([0.1])
    d_out, d_feat = rnd(n, 1, h, h, seed=16), rnd(n, c, h, h, seed=17, scale=0.01)
    zr, wr, br = z.double().requires_grad_(True), w.double().requires_grad_(True), b.double().requires_grad_(True)
    up = F.relu(zr)
    out = torch.sigmoid(F.conv2d(up, wr, br))
    ((out * d_out.double()).sum() + (up * d_feat.double()).sum()).backward()
    upb = to_blocked(F.relu(z))
    dz = torch.empty_like(upb)
    dw, dbo, dbu = torch.zeros(c, device="cuda"), torch.zeros(1, device="cuda"), torch.zeros(c, device="cuda")
    outg = torch.sigmoid(F.conv2d(F.relu(z), w, b)).cuda().contiguous()
    call("uncl_outc_feat_bwd", d_out.cuda(), outg, upb, upb.stride(0), to_blocked(d_feat), w.reshape(-1).cuda(), dz, dw, dbo, dbu,
         n, c, h * h)
    got = from_blocked(dz)
    assert rel(got, zr.grad) < 4e-3 and (got[z <= 0] == 0).all()
    assert rel(dw, wr.grad.reshape(-1)) < 1e-4 and rel(dbo, br.grad) < 1e-4
    assert rel(dbu, got.double().sum(dim=(0, 2, 3))) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_first_conv_weight_and_bias_gradient(dtype):
    n, c, h, w = 2, 32, 40, 45
    x, dz = rnd(n, 1, h, w, seed=20), rnd(n, c, h - 2, w - 2, seed=21)
    wt = torch.zeros(c, 1, 3, 3, dtype=torch.float64, requires_grad=True)
    b = torch.zeros(c, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double(), wt, b).backward(dz.double())
    dw, db = torch.zeros(9, c, device="cuda"), torch.zeros(c, device="cuda")
    call("uncl_conv_first_wgrad_bias", x.cuda(), to_blocked(dz, dtype), _lib.DTYPE_OF[dtype], dw, db, n, h, w, c)
    assert rel(dw, wt.grad.reshape(c, 9).t()) < 2e-6 and rel(db, b.grad) < 2e-6


def test_pack_and_unpack_gathers():
    g = torch.Generator().manual_seed(1)
    src = torch.randn(5000, generator=g).cuda()
    idx = torch.randint(0, 5000, (7777,), generator=g, dtype=torch.int32)
    idx[::11] = -1
    lo = torch.zeros(7777, dtype=torch.bool)
    lo[5::7] = True
    lo &= idx >= 0
    enc = (idx.long() | (lo.long() << 30)).to(torch.int32)
    enc[idx < 0] = -1
    out = torch.empty(7777, device="cuda", dtype=torch.bfloat16)
    call("uncl_pack_gather", src, enc.cuda(), out, 7777)
    v = src.cpu()[idx.clamp(min=0).long()]
    hi = v.to(torch.bfloat16)
    want = torch.where(idx < 0, torch.zeros_like(hi), torch.where(lo, (v - hi.float()).to(torch.bfloat16), hi))
    assert torch.equal(out.cpu(), want)
    dst = torch.ones(7777, device="cuda")
    call("uncl_unpack_add", src, idx.cuda(), dst, 7777)
    assert torch.equal(dst.cpu(), 1.0 + torch.where(idx < 0, torch.zeros_like(v), v))
    call("uncl_unpack_gather", src, idx.cuda(), dst, 7777)
    assert torch.equal(dst.cpu()[idx >= 0], v[idx >= 0])


@pytest.mark.parametrize("sel,use_ext", [((3, 1), False), ((4, 0), False), ((5, 6), True), ((1, 6), True)])
def test_self_nce_on_blocked_features(sel, use_ext):
    """uncl_nce_self_fwd / bwd against autograd through the oracle's nce: positive / negative = rows of the anchor tensor
    (or of `ext`, the data-parallel case), gradient of the selected rows summed over the batch."""
    b, chw, hw = 5, 8 * 96, 96
    fea = rnd(b, chw, seed=18).abs()
    ext = rnd(2, chw, seed=19).abs()
    fr, er = fea.double().requires_grad_(True), ext.double().requires_grad_(True)
    rows = torch.cat([fr, er])
    shape = (-1, chw // hw, hw, 1)
    pos = rows[sel[0]:sel[0] + 1].reshape(shape).expand(b, -1, -1, -1)
    neg = rows[sel[1]:sel[1] + 1].reshape(shape).expand(b, -1, -1, -1)
    loss = oracle.nce(fr.reshape(shape), pos, neg, 1.0, 1e-2)
    (loss * 0.7).backward()
    fb = fea.to(torch.bfloat16).cuda()
    eb = ext.to(torch.bfloat16).cuda() if use_ext else None
    s = torch.tensor(sel, dtype=torch.int64, device="cuda")
    logits, out = torch.empty(2 * b, device="cuda"), torch.empty((), device="cuda")
    call("uncl_nce_self_fwd", fb, s, eb, b, chw, hw, 1.0, 1e-2, logits, out)
    assert abs(out.item() - loss.item()) <= 1e-5 * abs(loss.item())
    d = torch.empty((b, chw), device="cuda")
    d_ext = torch.zeros((2, chw), device="cuda") if use_ext else None
    call("uncl_nce_self_bwd", fb, s, eb, b, chw, hw, 1.0, 1e-2, logits, torch.tensor(0.7, device="cuda"), d, F32, d_ext)
    assert rel(d, fr.grad) < 1e-5
    if use_ext:
        for j in range(2):
            if b + j in sel:
                assert rel(d_ext[j], er.grad[j]) < 1e-5
    d16 = torch.empty((b, chw), device="cuda", dtype=torch.bfloat16)
    call("uncl_nce_self_bwd", fb, s, eb, b, chw, hw, 1.0, 1e-2, logits, torch.tensor(0.7, device="cuda"), d16, BF16, d_ext)
    assert rel(d16.float(), fr.grad) < 4e-3


def test_flat_adam_matches_torch():
    g = torch.Generator().manual_seed(2)
    p0 = torch.randn(4097, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-3, betas=(0.5, 0.999))
    p = p0.clone().cuda()
    m, v, step = torch.zeros_like(p), torch.zeros_like(p), torch.zeros(1, device="cuda")
    for it in range(4):
        grad = torch.randn(4097, generator=g) * (it + 1)
        ref.grad = grad.clone()
        opt.step()
        call("uncl_adam_flat", p, grad.cuda(), m, v, 4097, 1e-3, 0.5, 0.999, 1e-8, step)
    assert step.item() == 4.0
    assert (p.cpu() - ref.detach()).abs().max().item() <= 2e-6
