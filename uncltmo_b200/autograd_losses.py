"""autograd Functions of the discriminator and the training losses (forward and backward are library kernels).
Upstream gradients reach the kernels as device scalars, so a training step never synchronises with the host."""
import ctypes

import torch
from torch.autograd import Function

from ._lib import BF16, call


def _f(t):
    return t.contiguous().float()


def _scalar(like):
    return torch.empty((), device=like.device, dtype=torch.float32)


class StructLossFn(Function):
    """StructLoss.forward (models/struct_loss.py:23-104)."""

    @staticmethod
    def forward(ctx, fake, hdr, weights):
        fake, hdr = _f(fake), _f(hdr)
        n, _, h, w = fake.shape
        wh = (ctypes.c_float * len(weights))(*weights)
        scratch = torch.empty(2 * n * (h // 2) * (w // 2) * 4 // 3 + 64, device=fake.device, dtype=torch.float32)
        out = _scalar(fake)
        call("uncl_struct_loss_fwd", fake, hdr, n, h, w, len(weights), ctypes.cast(wh, ctypes.c_void_p).value, out, scratch)
        ctx.save_for_backward(fake, hdr)
        ctx.weights = list(weights)
        return out

    @staticmethod
    def backward(ctx, g):
        fake, hdr = ctx.saved_tensors
        n, _, h, w = fake.shape
        wh = (ctypes.c_float * len(ctx.weights))(*ctx.weights)
        d = torch.empty_like(fake)
        scratch = torch.empty(int(6.5 * n * h * w) + 64, device=fake.device, dtype=torch.float32)
        call("uncl_struct_loss_bwd", fake, hdr, n, h, w, len(ctx.weights), ctypes.cast(wh, ctypes.c_void_p).value, _f(g), d,
             scratch)
        return d, None, None


class ContrastiveDFn(Function):
    """GanTrainer.contrastive_D_loss (GanTrainerImg.py:219-229)."""

    @staticmethod
    def forward(ctx, real_logits, fake_logits):
        r, f = _f(real_logits.reshape(-1)), _f(fake_logits.reshape(-1))
        out = _scalar(r)
        call("uncl_contrastive_d_loss", r, f, r.numel(), out)
        ctx.save_for_backward(r, f)
        ctx.shapes = (real_logits.shape, fake_logits.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        r, f = ctx.saved_tensors
        dr, df = torch.empty_like(r), torch.empty_like(f)
        call("uncl_contrastive_d_bwd", r, f, r.numel(), _f(g), dr, df)
        return dr.reshape(ctx.shapes[0]), df.reshape(ctx.shapes[1])


class NceFn(Function):
    """GanTrainer.nce, one positive / one negative, InfoNCE (GanTrainerImg.py:410-439).  pos / neg may have batch 1
    (broadcast over the batch, as infoNCE2 builds them)."""

    @staticmethod
    def forward(ctx, anchor, pos, neg, k, constant):
        a, p, n = _f(anchor), _f(pos), _f(neg)
        b, c, h, w = a.shape
        ps = 0 if p.shape[0] == 1 and b > 1 else c * h * w
        ns = 0 if n.shape[0] == 1 and b > 1 else c * h * w
        logits = torch.empty(2 * b, device=a.device, dtype=torch.float32)
        out = _scalar(a)
        call("uncl_nce_fwd", a, p, ps, n, ns, b, c, h * w, float(k), float(constant), logits, out)
        ctx.save_for_backward(a, p, n, logits)
        ctx.cfg = (ps, ns, float(k), float(constant))
        return out

    @staticmethod
    def backward(ctx, g):
        a, p, n, logits = ctx.saved_tensors
        ps, ns, k, constant = ctx.cfg
        b, c, h, w = a.shape
        da = torch.empty_like(a)
        dp = torch.empty_like(p) if ctx.needs_input_grad[1] else None
        dn = torch.empty_like(n) if ctx.needs_input_grad[2] else None
        call("uncl_nce_bwd", a, p, ps, n, ns, b, c, h * w, k, constant, logits, _f(g), da, dp, dn)
        return da, dp, dn, None, None


class NceSelfFn(Function):
    """GanTrainer.nce as infoNCE2 calls it (GanTrainerImg.py:384-439) on a feature tensor in ANY element order (the
    similarity is a sum over all elements): the C8-blocked bf16 `up_x` of the bf16 training path is read as it is.
    sel: device int64 [2] = (row of the positive, row of the negative).  Rows < B are rows of `fea` itself; rows B, B+1
    are the two rows of `ext` (bf16 [2, ...], data-parallel training: the globally selected samples, see
    uncltmo_b200.dist.BroadcastRowsFn), whose gradient is returned for the owning rank to add up."""

    @staticmethod
    def forward(ctx, fea, sel, hw, k, constant, ext=None):
        fea = fea.contiguous()
        if fea.dtype != torch.bfloat16:
            raise TypeError("NceSelfFn reads bf16 features")
        b = fea.shape[0]
        chw = fea.numel() // b
        if ext is not None:
            ext = ext.contiguous()
        logits = torch.empty(2 * b, device=fea.device, dtype=torch.float32)
        out = _scalar(fea)
        call("uncl_nce_self_fwd", fea, sel, ext, b, chw, hw, float(k), float(constant), logits, out)
        ctx.save_for_backward(fea, sel, logits, ext)
        ctx.cfg = (hw, float(k), float(constant))
        return out

    @staticmethod
    def backward(ctx, g):
        fea, sel, logits, ext = ctx.saved_tensors
        hw, k, constant = ctx.cfg
        b = fea.shape[0]
        d = torch.empty_like(fea)
        d_ext = torch.zeros(ext.shape, device=fea.device, dtype=torch.float32) if ext is not None else None
        call("uncl_nce_self_bwd", fea, sel, ext, b, fea.numel() // b, hw, k, constant, logits, _f(g), d, BF16, d_ext)
        return d, None, None, None, None, d_ext


class PlaneMeanContrastFn(Function):
    """x [..., H, W] -> (per-plane mean, per-plane mean local variance under the 11x11 gaussian)."""

    @staticmethod
    def forward(ctx, x):
        x = _f(x)
        h, w = x.shape[-2], x.shape[-1]
        m = x.numel() // (h * w)
        mean = torch.empty(x.shape[:-2], device=x.device, dtype=torch.float32)
        con = torch.empty(x.shape[:-2], device=x.device, dtype=torch.float32)
        scratch = torch.empty(2 * m, device=x.device, dtype=torch.float32)
        call("uncl_plane_mean_contrast", x, h * w, m, h, w, mean, con, scratch)
        ctx.save_for_backward(x)
        return mean, con

    @staticmethod
    def backward(ctx, d_mean, d_con):
        (x,) = ctx.saved_tensors
        h, w = x.shape[-2], x.shape[-1]
        m = x.numel() // (h * w)
        dx = torch.empty_like(x)
        mu = torch.empty(m * (h - 10) * (w - 10), device=x.device, dtype=torch.float32)
        call("uncl_plane_mean_contrast_bwd", x, m, h, w, _f(d_mean) if d_mean is not None else None,
             _f(d_con) if d_con is not None else None, dx, mu)
        return dx


class L1MeanFn(Function):
    """nn.L1Loss of two small tensors."""

    @staticmethod
    def forward(ctx, a, b):
        ctx.shapes = (a.shape, b.shape)
        a, b = _f(a).reshape(-1), _f(b).reshape(-1)
        out = _scalar(a)
        call("uncl_l1_mean", a, b, a.numel(), out)
        ctx.save_for_backward(a, b)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        da = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        db = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        call("uncl_l1_mean_bwd", a, b, a.numel(), _f(g), da, db)
        return (da.reshape(ctx.shapes[0]) if da is not None else None,
                db.reshape(ctx.shapes[1]) if db is not None else None)


class TVFn(Function):
    """L_TV (GanTrainer.py:669-682)."""

    @staticmethod
    def forward(ctx, x):
        x = _f(x)
        b, c, h, w = x.shape
        scratch = torch.empty(2, device=x.device, dtype=torch.float32)
        out = _scalar(x)
        call("uncl_tv_loss", x, b, c, h, w, scratch, out)
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        b, c, h, w = x.shape
        dx = torch.empty_like(x)
        call("uncl_tv_bwd", x, b, c, h, w, _f(g), dx)
        return dx


class DiscFn(Function):
    """SimpleDiscriminator trunk + tail (models/Discriminator.py:98-122): x -> (logits [N,1], fea map [N,1,62,62])."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, w3, b3, wt):
        x = _f(x)
        n = x.shape[0]
        dev = x.device
        h1 = torch.empty((n, 16, 127, 127), device=dev, dtype=torch.float32)
        a2 = torch.empty((n, 32, 62, 62), device=dev, dtype=torch.float32)
        fea = torch.empty((n, 1, 62, 62), device=dev, dtype=torch.float32)
        logits = torch.empty((n, 1), device=dev, dtype=torch.float32)
        params = [_f(t.detach()) for t in (w1, b1, w2, b2, w3, b3, wt)]
        call("uncl_disc_forward", x, params[0], params[1], params[2], params[3], params[4], params[5], params[6], h1, a2, fea,
             logits, n, 256, 256)
        ctx.save_for_backward(x, h1, a2, fea, params[0], params[2], params[4], params[6])
        return logits, fea

    @staticmethod
    def backward(ctx, d_logits, d_fea):
        x, h1, a2, fea, w1, w2, w3, wt = ctx.saved_tensors
        n = x.shape[0]
        dev = x.device
        dfe = _f(d_fea).clone() if d_fea is not None else torch.zeros_like(fea)
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        z = lambda *s: torch.zeros(s, device=dev, dtype=torch.float32)  # noqa: E731
        want_w = any(ctx.needs_input_grad[1:])
        dw3, db3, dwt = z(1, 32, 1, 1), z(1), z(1, 62 * 62)
        # parameters frozen by the caller (the generator step back-propagates THROUGH the discriminator): the two
        # convolutions' weight gradients - a third of the backward's time - are not computed
        dw1, db1, dw2, db2 = (z(16, 1, 4, 4), z(16), z(32, 16, 4, 4), z(32)) if want_w else (None, None, None, None)
        scratch = torch.empty(n * (32 * 62 * 62 + 16 * 127 * 127), device=dev, dtype=torch.float32)
        call("uncl_disc_backward", x, h1, a2, fea, w1, w2, w3, wt, _f(d_logits) if d_logits is not None else None, dfe, dx,
             dw1, db1, dw2, db2, dw3, db3, dwt, scratch, n)
        if not want_w:
            return dx, None, None, None, None, None, None, None
        return dx, dw1, db1, dw2, db2, dw3, db3, dwt


def tmqi_naturalness(x):
    """TMQI statistical naturalness N per image of x [M,1,H,W] in [0,1] (no gradient; TMQI.py:210-242)."""
    x = _f(x.detach())
    m, h, w = x.shape[0], x.shape[-2], x.shape[-1]
    out = torch.empty(m, device=x.device, dtype=torch.float32)
    scratch = torch.empty(2 * m, device=x.device, dtype=torch.float32)
    call("uncl_tmqi_naturalness", x, m, h, w, scratch, out)
    return out
