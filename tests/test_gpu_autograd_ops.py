"""Unit parity of each autograd Function (forward + backward kernels) against torch autograd on CPU (float64)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from uncltmo_b200 import autograd as A

pytestmark = pytest.mark.gpu


def to_blocked(x):  # NCHW -> [N, C/8, H, W, 8]
    n, c, h, w = x.shape
    return x.reshape(n, c // 8, 8, h, w).permute(0, 1, 3, 4, 2).contiguous()


def from_blocked(x):
    n, cb, h, w, _ = x.shape
    return x.permute(0, 1, 4, 2, 3).reshape(n, cb * 8, h, w)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def rnd(*shape, seed=0, scale=1.0):
    return torch.from_numpy((np.random.default_rng(seed).standard_normal(shape) * scale).astype(np.float32))


def leaf(t, dev=None):
    t = t.clone().to(dev) if dev else t.clone().double()
    return t.requires_grad_(True)


@pytest.mark.parametrize("ci,co,h,transposed,relu", [(32, 32, 20, False, True), (32, 64, 37, False, True),
                                                    (64, 32, 13, True, True), (128, 128, 9, True, False),
                                                    (256, 256, 12, False, True), (32, 32, 33, True, True)])
def test_conv3x3(ci, co, h, transposed, relu):
    x = rnd(2, ci, h, h + 3, seed=1)
    w = rnd(*((ci, co, 3, 3) if transposed else (co, ci, 3, 3)), seed=2, scale=(9 * ci) ** -0.5)
    b = rnd(co, seed=3, scale=0.1)
    xr, wr, br = leaf(x), leaf(w), leaf(b)
    yr = F.conv_transpose2d(xr, wr, br) if transposed else F.conv2d(xr, wr, br)
    yr = F.relu(yr) if relu else yr
    g = rnd(*yr.shape, seed=4)
    yr.backward(g.double())
    xb, wc, bc = leaf(to_blocked(x), "cuda"), leaf(w, "cuda"), leaf(b, "cuda")
    y = A.Conv3x3.apply(xb, wc, bc, transposed, relu)
    y.backward(to_blocked(g).cuda())
    assert rel(from_blocked(y), yr) < 1e-5
    assert rel(from_blocked(xb.grad), xr.grad) < 1e-5
    assert rel(wc.grad, wr.grad) < 1e-5 and rel(bc.grad, br.grad) < 1e-5


@pytest.mark.parametrize("ci,co,h,transposed", [(32, 32, 70, False), (32, 64, 37, False), (64, 128, 30, False),
                                               (128, 32, 66, True), (256, 256, 12, False), (256, 256, 10, True),
                                               (1024, 128, 24, True), (512, 64, 57, True), (128, 128, 59, False)])
def test_conv3x3_tensor_core_path(ci, co, h, transposed):
    """tc=True: tcgen05 forward, data gradient and weight gradient (bf16 operands, fp32 accumulation) against float64.
    Inputs are pre-rounded to bf16 so that only the accumulation order differs."""
    x = rnd(2, ci, h, h + 2, seed=1).to(torch.bfloat16).float()
    w = rnd(*((ci, co, 3, 3) if transposed else (co, ci, 3, 3)), seed=2, scale=(9 * ci) ** -0.5).to(torch.bfloat16).float()
    b = rnd(co, seed=3, scale=0.1)
    xr, wr, br = leaf(x), leaf(w), leaf(b)
    # no ReLU here: a pre-activation within fp32 rounding of 0 would flip its mask between the two sides and change the
    # 3x3 neighbourhood of that one pixel (the mask itself is covered by test_conv3x3)
    yr = F.conv_transpose2d(xr, wr, br) if transposed else F.conv2d(xr, wr, br)
    g = rnd(*yr.shape, seed=4).to(torch.bfloat16).float()
    yr.backward(g.double())
    xb, wc, bc = leaf(to_blocked(x), "cuda"), leaf(w, "cuda"), leaf(b, "cuda")
    y = A.Conv3x3.apply(xb, wc, bc, transposed, False, True)
    y.backward(to_blocked(g).cuda())
    assert rel(from_blocked(y), yr) < 1e-5
    assert rel(from_blocked(xb.grad), xr.grad) < 1e-5      # dz is exactly representable (g * mask), weights bf16
    assert rel(wc.grad, wr.grad) < 1e-5 and rel(bc.grad, br.grad) < 1e-5


@pytest.mark.parametrize("ci,co,h,transposed", [(32, 32, 40, False), (64, 128, 30, False), (128, 32, 66, True),
                                               (256, 256, 12, False), (512, 64, 29, True)])
def test_conv3x3_exact_split_path(ci, co, h, transposed):
    """tc="split" (precision 'fp32_tc'): forward, data gradient and weight gradient as three-term bf16 split GEMMs on the
    tensor cores with UNROUNDED fp32 inputs against float64 - ~2^-16 per product, i.e. fp32-grade results (the bf16 path
    on the same inputs would sit at ~4e-3).  No ReLU: a pre-activation within 1e-5 of 0 would take a different side of the
    mask on the two sides (the mask itself is covered by test_conv3x3)."""
    relu = False
    x = rnd(2, ci, h, h + 2, seed=1)
    w = rnd(*((ci, co, 3, 3) if transposed else (co, ci, 3, 3)), seed=2, scale=(9 * ci) ** -0.5)
    b = rnd(co, seed=3, scale=0.1)
    xr, wr, br = leaf(x), leaf(w), leaf(b)
    yr = F.conv_transpose2d(xr, wr, br) if transposed else F.conv2d(xr, wr, br)
    yr = F.relu(yr) if relu else yr
    g = rnd(*yr.shape, seed=4)
    yr.backward(g.double())
    xb, wc, bc = leaf(to_blocked(x), "cuda"), leaf(w, "cuda"), leaf(b, "cuda")
    y = A.Conv3x3.apply(xb, wc, bc, transposed, relu, "split")
    y.backward(to_blocked(g).cuda())
    assert rel(from_blocked(y), yr) < 2e-5
    assert rel(from_blocked(xb.grad), xr.grad) < 2e-5
    assert rel(wc.grad, wr.grad) < 2e-5 and rel(bc.grad, br.grad) < 1e-5


def test_conv_first():
    x = rnd(2, 1, 40, 52, seed=1)
    w, b = rnd(32, 1, 3, 3, seed=2, scale=0.3), rnd(32, seed=3, scale=0.1)
    wr, br = leaf(w), leaf(b)
    yr = F.relu(F.conv2d(x.double(), wr, br))
    g = rnd(*yr.shape, seed=4)
    yr.backward(g.double())
    wc, bc = leaf(w, "cuda"), leaf(b, "cuda")
    y = A.ConvFirst.apply(x.cuda(), wc, bc)
    y.backward(to_blocked(g).cuda())
    assert rel(from_blocked(y), yr) < 1e-5 and rel(wc.grad, wr.grad) < 1e-5 and rel(bc.grad, br.grad) < 1e-5


@pytest.mark.parametrize("h,w", [(24, 24), (57, 61), (7, 10)])
def test_maxpool(h, w):
    x = rnd(2, 16, h, w, seed=5)
    x[0, :, :4, :4] = 0.25  # ties: the first maximum takes the gradient
    xr = leaf(x)
    yr = F.max_pool2d(xr, 2)
    g = rnd(*yr.shape, seed=6)
    yr.backward(g.double())
    xb = leaf(to_blocked(x), "cuda")
    y = A.MaxPool2.apply(xb)
    y.backward(to_blocked(g).cuda())
    assert rel(from_blocked(y), yr) == 0.0 and rel(from_blocked(xb.grad), xr.grad) < 1e-7


@pytest.mark.parametrize("tc", [False, True])
@pytest.mark.parametrize("c,h,h2", [(32, 12, 24), (64, 28, 57), (256, 6, 12)])
def test_convT2x2(c, h, h2, tc):
    x = rnd(2, c, h, h, seed=7)
    w, b = rnd(c, c, 2, 2, seed=8, scale=c ** -0.5), rnd(c, seed=9, scale=0.1)
    xr, wr, br = leaf(x), leaf(w), leaf(b)
    yr = F.conv_transpose2d(xr, wr, br, stride=2)
    d = h2 - 2 * h
    if d:
        yr = F.pad(yr, (d // 2, d - d // 2, d // 2, d - d // 2), mode="replicate")
    g = rnd(*yr.shape, seed=10)
    yr.backward(g.double())
    xb, wc, bc = leaf(to_blocked(x), "cuda"), leaf(w, "cuda"), leaf(b, "cuda")
    y = A.ConvT2x2.apply(xb, wc, bc, h2, h2, tc)
    y.backward(to_blocked(g).cuda())
    tol = 6e-3 if tc else 1e-5   # tc: bf16-rounded operands (2^-9 each), fp32 accumulation
    assert rel(from_blocked(y), yr) < tol
    assert rel(from_blocked(xb.grad), xr.grad) < tol
    assert rel(wc.grad, wr.grad) < 1e-5 and rel(bc.grad, br.grad) < 1e-5   # weight / bias gradients stay fp32


def test_skip_concat():
    x2 = rnd(2, 32, 9, 11, seed=11).abs()
    x2[0, :, :2] = 0.0  # post-ReLU zeros: d sqrt(x+1e-8) = 5000
    x1 = rnd(2, 32, 9, 11, seed=12)
    ar, br = leaf(x2), leaf(x1)
    yr = torch.cat([ar, br, ar * ar, torch.pow(ar + 1e-8, 0.5)], dim=1)
    g = rnd(*yr.shape, seed=13)
    yr.backward(g.double())
    a, b = leaf(to_blocked(x2), "cuda"), leaf(to_blocked(x1), "cuda")
    y = A.SkipConcat.apply(a, b)
    y.backward(to_blocked(g).cuda())
    assert rel(from_blocked(y), yr) < 1e-6
    assert rel(from_blocked(a.grad), ar.grad) < 1e-5 and rel(from_blocked(b.grad), br.grad) < 1e-7


def blocked144(x):  # [N,C,12,12] -> [N, C/8, 144, 8]
    return to_blocked(x).reshape(x.shape[0], x.shape[1] // 8, 144, 8)


@pytest.mark.parametrize("tc", [False, True, "split"])
@pytest.mark.parametrize("ci,co,groups,gelu,with_res", [(256, 256, 1, False, False), (512, 512, 4, True, False),
                                                       (512, 256, 1, False, True), (256, 256, 1, True, False)])
def test_pw_conv(ci, co, groups, gelu, with_res, tc):
    if tc == "split" and groups != 1:
        pytest.skip("the split forward is built for ungrouped 1x1 convs (fc1)")
    x = rnd(3, ci, 12, 12, seed=14)
    w, b = rnd(co, ci // groups, 1, 1, seed=15, scale=(ci // groups) ** -0.5), rnd(co, seed=16, scale=0.1)
    res = rnd(3, co, 12, 12, seed=17) if with_res else None
    scale = torch.tensor([1 / 0.95, 0.0, 1 / 0.95]) if with_res else None
    xr, wr, br = leaf(x), leaf(w), leaf(b)
    rr = leaf(res) if with_res else None
    yr = F.conv2d(xr, wr, br, groups=groups)
    yr = F.gelu(yr) if gelu else yr
    if with_res:
        yr = yr * scale.double().view(3, 1, 1, 1) + rr
    g = rnd(*yr.shape, seed=18)
    yr.backward(g.double())
    xb, wc, bc = leaf(blocked144(x), "cuda"), leaf(w, "cuda"), leaf(b, "cuda")
    rb = leaf(blocked144(res), "cuda") if with_res else None
    y = A.PwConv.apply(xb, wc, bc, rb, scale.cuda() if with_res else None, groups, gelu, tc)
    y.backward(blocked144(g).cuda())
    unb = lambda t: from_blocked(t.reshape(t.shape[0], t.shape[1], 12, 12, 8))  # noqa: E731
    tol = 6e-3 if tc else 1e-5   # tc: bf16-rounded operands, fp32 accumulation
    assert rel(unb(y), yr) < (2e-5 if tc == "split" else tol)     # three-term split: ~2^-16 per product
    assert rel(unb(xb.grad), xr.grad) < tol
    # tc: the weight gradient is a tensor-core GEMM on bf16-rounded x and dz; the bias gradient stays fp32
    assert rel(wc.grad, wr.grad) < tol and rel(bc.grad, br.grad) < (tol if gelu else 1e-5)
    if with_res:
        assert rel(unb(rb.grad), rr.grad) < 1e-7


def test_knn_aggregate_and_addpos():
    import oracle
    from uncltmo_b200.weights import relative_pos_table
    y = rnd(2, 256, 12, 12, seed=19)
    pos = rnd(1, 256, 12, 12, seed=20, scale=0.1)
    relpos = relative_pos_table()
    yr, pr = leaf(y), leaf(pos)
    x0 = (yr + pr).reshape(2, 256, 144)
    idx = oracle.knn_indices(x0.detach().float(), relpos)
    yj = torch.gather(x0.unsqueeze(-1).expand(2, 256, 144, 9), 2, idx.unsqueeze(1).expand(2, 256, 144, 9))
    agg = (yj - x0.unsqueeze(-1)).max(dim=-1)[0]
    zr = torch.stack([x0, agg], dim=2).reshape(2, 512, 12, 12)
    g = rnd(*zr.shape, seed=21)
    zr.backward(g.double())
    yb, pc = leaf(to_blocked(y), "cuda"), leaf(pos, "cuda")
    x0c = A.AddPos.apply(yb, pc)
    z = A.KnnAggregate.apply(x0c, relpos.reshape(144, 144).cuda())
    z.backward(blocked144(g).cuda())
    unb = lambda t: from_blocked(t.reshape(t.shape[0], t.shape[1], 12, 12, 8))  # noqa: E731
    assert rel(unb(z), zr) < 1e-6
    assert rel(from_blocked(yb.grad), yr.grad) < 1e-6 and rel(pc.grad, pr.grad) < 1e-6


def test_outc_sigmoid_and_layout():
    up = rnd(2, 32, 20, 24, seed=22)
    w, b = rnd(1, 32, 1, 1, seed=23, scale=0.2), rnd(1, seed=24)
    ur, wr, br = leaf(up), leaf(w), leaf(b)
    outr = torch.sigmoid(F.conv2d(ur, wr, br))
    g, gf = rnd(*outr.shape, seed=25), rnd(*up.shape, seed=26)
    (outr * g.double()).sum().backward(retain_graph=True)
    (ur * gf.double()).sum().backward()
    ub, wc, bc = leaf(to_blocked(up), "cuda"), leaf(w, "cuda"), leaf(b, "cuda")
    out = A.OutcSigmoid.apply(ub, wc, bc)
    feats = A.BlockedToNCHW.apply(ub)
    ((out * g.cuda()).sum() + (feats * gf.cuda()).sum()).backward()
    assert rel(out, outr) < 1e-6 and rel(feats, up) == 0.0
    assert rel(from_blocked(ub.grad), ur.grad) < 1e-6
    assert rel(wc.grad, wr.grad) < 1e-5 and rel(bc.grad, br.grad) < 1e-5


@pytest.mark.parametrize("c,h,r", [(32, 20, 1), (64, 13, 2), (128, 7, 4), (256, 12, 8)])
def test_splice_channels(c, h, r):
    """Video hand-over cat(prev[:, :r], cur[:, r:]) (Unet.py:244, 270): exact copy forward, gradient split backward."""
    cur, prev = rnd(2, c, h, h, seed=1), rnd(2, c, h, h, seed=2)
    cr, pr = leaf(cur), leaf(prev)
    yr = torch.cat((pr[:, :r], cr[:, r:]), 1)
    g = rnd(*yr.shape, seed=3)
    yr.backward(g.double())
    cb, pb = leaf(to_blocked(cur), "cuda"), leaf(to_blocked(prev), "cuda")
    y = A.SpliceChannels.apply(cb, pb, r)
    y.backward(to_blocked(g).cuda())
    assert torch.equal(from_blocked(y).cpu().double(), yr.detach())
    assert torch.equal(from_blocked(cb.grad).cpu().double(), cr.grad)
    assert torch.equal(from_blocked(pb.grad).cpu().double(), pr.grad)
