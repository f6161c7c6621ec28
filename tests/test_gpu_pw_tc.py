"""The pointwise (1x1, grouped) convolution as a tensor-core GEMM (uncl_pw_conv_tc) against torch on the same
bf16-rounded operands (float64 accumulation): fp32 output within 1e-5, bf16 output within bf16 rounding."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from uncltmo_b200 import _lib, packing
from uncltmo_b200._lib import call

pytestmark = pytest.mark.gpu


def to_blocked(x):
    n, c, h, w = x.shape
    return x.reshape(n, c // 8, 8, h, w).permute(0, 1, 3, 4, 2).contiguous()


def from_blocked(x):
    n, cb, h, w, _ = x.shape
    return x.permute(0, 1, 4, 2, 3).reshape(n, cb * 8, h, w)


def rnd(*shape, seed=0, scale=1.0):
    return torch.from_numpy((np.random.default_rng(seed).standard_normal(shape) * scale).astype(np.float32))


@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():
        yield


@pytest.mark.parametrize("n,ci,co,groups,h,w,act,res,scale,out_dtype", [
    (3, 512, 512, 4, 12, 12, "gelu", None, False, torch.bfloat16),     # graph conv
    (3, 512, 256, 1, 12, 12, "none", "f32", True, torch.bfloat16),     # Grapher fc2 + residual + DropPath scale
    (2, 256, 256, 1, 12, 12, "gelu", None, False, torch.float32),      # FFN fc1
    (2, 256, 256, 1, 12, 12, "none", "bf16", True, torch.bfloat16),    # FFN fc2
    (2, 128, 32, 1, 126, 126, "none", None, False, torch.float32),     # up-conv data gradient (4C -> C), 126^2
    (1, 512, 128, 1, 28, 28, "relu", None, False, torch.float32),
    (2, 1024, 256, 1, 12, 12, "none", None, False, torch.float32),
    (1, 64, 64, 2, 61, 61, "none", "f32", False, torch.float32),
])
def test_pw_conv_tc(n, ci, co, groups, h, w, act, res, scale, out_dtype):
    x = rnd(n, ci, h, w, seed=1).bfloat16()
    wt = rnd(co, ci // groups, 1, 1, seed=2, scale=(ci // groups) ** -0.5).bfloat16().float()
    b = rnd(co, seed=3, scale=0.1)
    r = rnd(n, co, h, w, seed=4) if res else None
    if res == "bf16":
        r = r.bfloat16().float()
    sc = torch.tensor([1.0 / 0.95, 0.0, 1.0][:n]) if scale else None
    want = F.conv2d(x.double(), wt.double(), b.double(), groups=groups)
    want = {"gelu": F.gelu, "relu": F.relu, "none": lambda t: t}[act](want)
    if sc is not None:
        want = want * sc.double().view(n, 1, 1, 1)
    if r is not None:
        want = want + r.double()

    xb = to_blocked(x).cuda()
    out = torch.empty((n, co // 8, h, w, 8), device="cuda", dtype=out_dtype)
    rb = None
    if r is not None:
        rb = to_blocked(r).cuda().to(torch.bfloat16 if res == "bf16" else torch.float32)
    code = {"gelu": _lib.ACT_GELU, "relu": _lib.ACT_RELU, "none": _lib.ACT_NONE}[act]
    call("uncl_pw_conv_tc", xb, xb.stride(0), packing.pointwise_tc(wt.cuda(), groups), b.cuda(), rb,
         rb.stride(0) if rb is not None else 0, _lib.DTYPE_OF[rb.dtype] if rb is not None else _lib.F32,
         sc.cuda() if sc is not None else None, out, out.stride(0), _lib.DTYPE_OF[out_dtype], n, ci, co, groups, h, w, code)
    got = from_blocked(out.float()).cpu().double()
    err = ((got - want).norm() / want.norm()).item()
    assert err < (3e-3 if out_dtype == torch.bfloat16 else 1e-5), err
    if out_dtype == torch.bfloat16:
        assert torch.equal(got.float(), want.float().bfloat16().float()) or (got - want).abs().max() <= want.abs().max() * 2 ** -7


def test_pw_conv_tc_rejects_unsupported():
    x = torch.zeros((1, 2, 12, 12, 8), device="cuda", dtype=torch.bfloat16)
    out = torch.zeros((1, 4, 12, 12, 8), device="cuda")
    with pytest.raises(RuntimeError):
        call("uncl_pw_conv_tc", x, x.stride(0), x, None, None, 0, _lib.F32, None, out, out.stride(0), _lib.F32, 1, 16, 24, 1,
             12, 12, _lib.ACT_NONE)
