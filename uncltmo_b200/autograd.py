"""torch.autograd.Function wrappers: every forward AND backward is a kernel of libuncltmo_b200.so.

Autograd is used only as the graph engine (ordering, gradient accumulation into `.grad`); no torch operator computes
on the hot path.  Training runs on the fp32 path: activations are C8-blocked fp32 tensors `[N, C/8, H, W, 8]`
(`[N, C/8, 144, 8]` inside the graph block); parameters keep the reference layouts.
"""
import torch
from torch.autograd import Function

from . import _lib, packing
from ._lib import ACT_GELU, ACT_NONE, ACT_RELU, BF16, F32, call


def _zeros(shape, like):
    return torch.zeros(shape, device=like.device, dtype=torch.float32)


def _empty(shape, like):
    return torch.empty(shape, device=like.device, dtype=torch.float32)


class ConvFirst(Function):
    """inc.conv.conv: Conv2d(1, C, 3) + ReLU on a plain [N,1,H,W] image (unet_parts.py:57-87)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x = x.contiguous().float()
        n, _, h, w = x.shape
        c = weight.shape[0]
        y = _empty((n, c // 8, h - 2, w - 2, 8), x)
        call("uncl_conv_first", x, packing.conv_first(weight.detach()), bias.detach().float().contiguous(), y, y.stride(0),
             n, h, w, c, ACT_RELU, F32)
        ctx.save_for_backward(x, y)
        ctx.c = c
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        n, _, h, w = x.shape
        c = ctx.c
        dz = dy.contiguous().clone()
        db = _zeros(c, x)
        call("uncl_relu_bwd_bias", dz, y, y.stride(0), db, n, c, (h - 2) * (w - 2), 1)
        dw = _zeros((9, c), x)
        call("uncl_conv_first_wgrad", x, dz, dw, n, h, w, c)
        return None, dw.t().reshape(c, 1, 3, 3), db


def _to_bf16(t):
    o = torch.empty(t.shape, device=t.device, dtype=torch.bfloat16)
    call("uncl_convert", t, F32, o, _lib.BF16, t.numel())
    return o


def _split3(x, n, c, hw):
    """fp32 blocked [N][C/8][HW][8] -> bf16 [N][3C/8][HW][8] = [hi | hi | lo] (uncl_split_bf16)."""
    xs = torch.empty((n, 3 * c // 8) + tuple(x.shape[2:]), device=x.device, dtype=torch.bfloat16)
    call("uncl_split_bf16", x, x.stride(0), xs, xs.stride(0), n, c, hw)
    return xs


class Conv3x3(Function):
    """nn.Conv2d(k=3, valid) or nn.ConvTranspose2d(k=3, s=1, p=0) (+ReLU).  unet_parts.py:57-87, 126-141, 183-193.

    tc=False: fp32 CUDA-core kernels.  tc=True ("mixed" training): forward and data gradient run on the tcgen05
    kernel with bf16-rounded operands and fp32 accumulation / fp32 outputs; the weight gradient stays fp32."""

    @staticmethod
    def forward(ctx, x, weight, bias, transposed, relu, tc=False):
        w9 = packing.conv3x3_taps(weight.detach(), transposed)  # [9][C_in][C_out]
        n, cb, h, w, _ = x.shape
        ci, co = w9.shape[1], w9.shape[2]
        pad = 2 if transposed else 0
        ho, wo = h + 2 * pad - 2, w + 2 * pad - 2
        x = x.contiguous()
        y = _empty((n, co // 8, ho, wo, 8), x)
        b = bias.detach().float().contiguous()
        act = ACT_RELU if relu else ACT_NONE
        if tc == "split":
            # exact path on the tensor cores: [x_hi | x_hi | x_lo] against [w_hi ; w_lo ; w_hi], fp32 accumulation (~2^-16)
            xs = _split3(x, n, ci, h * w)
            call("uncl_conv3x3_tc", xs, xs.stride(0), packing.conv3x3_tc_split(w9), b, y, y.stride(0), F32, n, 3 * ci, h, w, co,
                 pad, act, 0, 0, None, None, None, None)
            x = xs   # the backward reads the hi / lo thirds of the split copy (weight gradient on the tensor cores)
        elif tc:
            xb = _to_bf16(x)
            call("uncl_conv3x3_tc", xb, xb.stride(0), packing.conv3x3_tc(w9), b, y, y.stride(0), F32, n, ci, h, w, co, pad,
                 act, 0, 0, None, None, None, None)
            x = xb   # the backward only needs the bf16 copy (weight gradient on the tensor cores)
        else:
            call("uncl_conv3x3_simt", x, x.stride(0), w9, b, y, y.stride(0), n, ci, h, w, co, pad, act, 0, F32)
        ctx.save_for_backward(x, y, w9)
        ctx.cfg = (transposed, relu, pad, tc)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, w9 = ctx.saved_tensors
        transposed, relu, pad, tc = ctx.cfg
        n, _, h, w, _ = x.shape
        ci, co = w9.shape[1], w9.shape[2]
        ho, wo = y.shape[2], y.shape[3]
        dy = dy.contiguous()
        db = _zeros(co, dy)
        split = tc == "split"
        # one pass: ReLU mask + bias gradient + the operand the gradient GEMMs read (bf16 on the tensor-core path)
        dz = torch.empty(dy.shape, device=dy.device, dtype=torch.bfloat16 if (tc and not split) else torch.float32)
        call("uncl_relu_bwd_bias_out", dy, y, y.stride(0), dz, BF16 if (tc and not split) else F32, db, n, co, ho * wo,
             1 if relu else 0)
        dx = None
        dzb = dz
        if split:
            dzs = _split3(dz, n, co, ho * wo)     # [dz_hi | dz_hi | dz_lo]
            if ctx.needs_input_grad[0]:
                wt = w9.flip(0).transpose(1, 2).contiguous()
                dx = torch.empty((n, ci // 8, h, w, 8), device=dy.device, dtype=torch.float32)
                call("uncl_conv3x3_tc", dzs, dzs.stride(0), packing.conv3x3_tc_split(wt), _zeros(ci, dy), dx, dx.stride(0), F32,
                     n, 3 * co, ho, wo, ci, 2 - pad, ACT_NONE, 0, 0, None, None, None, None)
            dw9 = _zeros((9, ci, co), dy)
            cb, cob = ci // 8, co // 8
            x_hi, x_lo, z_hi, z_lo = x[:, :cb], x[:, 2 * cb:], dzs[:, :cob], dzs[:, 2 * cob:]
            for xa, za in ((x_hi, z_hi), (x_hi, z_lo), (x_lo, z_hi)):   # three-term split, accumulated by the kernel's atomics
                call("uncl_conv3x3_wgrad_tc_strided", xa, x.stride(0), za, dzs.stride(0), dw9, n, ci, h, w, co, pad)
            if transposed:
                dw = dw9.reshape(3, 3, ci, co).permute(2, 3, 0, 1).flip(2, 3)
            else:
                dw = dw9.reshape(3, 3, ci, co).permute(3, 2, 0, 1)
            return dx, dw.contiguous(), db, None, None, None
        if ctx.needs_input_grad[0]:
            # dgrad of a correlation with pad p = correlation of dz with pad 2-p and the taps reversed / transposed
            wt = w9.flip(0).transpose(1, 2).contiguous()
            dx = _empty(x.shape, dy)
            if tc:
                call("uncl_conv3x3_tc", dzb, dzb.stride(0), packing.conv3x3_tc(wt), _zeros(ci, dy), dx, dx.stride(0), F32, n,
                     co, ho, wo, ci, 2 - pad, ACT_NONE, 0, 0, None, None, None, None)
            else:
                call("uncl_conv3x3_simt", dz, dz.stride(0), wt, _zeros(ci, dy), dx, dx.stride(0), n, co, ho, wo, ci, 2 - pad,
                     ACT_NONE, 0, F32)
        dw9 = _zeros((9, ci, co), dy)
        if tc:
            call("uncl_conv3x3_wgrad_tc", x, x.stride(0), dzb, dw9, n, ci, h, w, co, pad)
        else:
            call("uncl_conv3x3_wgrad", x, x.stride(0), dz, dw9, n, ci, h, w, co, pad)
        if transposed:   # w9[t] = W[:, :, 2-ky, 2-kx]  (W is [C_in, C_out, 3, 3])
            dw = dw9.reshape(3, 3, ci, co).permute(2, 3, 0, 1).flip(2, 3)
        else:            # w9[t] = W[co, ci, ky, kx]
            dw = dw9.reshape(3, 3, ci, co).permute(3, 2, 0, 1)
        return dx, dw.contiguous(), db, None, None, None


class SpliceChannels(Function):
    """Video recurrence: cat(prev[:, :r], cur[:, r:]) (Unet.py:244, 270); the gradient flows to both (prev is not
    detached in the reference)."""

    @staticmethod
    def forward(ctx, cur, prev, r):
        out = cur.contiguous().clone()
        prev = prev.contiguous()
        n = out.shape[0]
        hw = out.numel() // (n * out.shape[1] * 8)
        call("uncl_splice_channels", out, out.stride(0), prev, prev.stride(0), r, n, hw, F32)
        ctx.r = r
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = dout.contiguous()
        n = dout.shape[0]
        hw = dout.numel() // (n * dout.shape[1] * 8)
        d_cur = dout.clone()
        zero = torch.zeros((n, 1) + tuple(dout.shape[2:]), device=dout.device, dtype=torch.float32)
        call("uncl_splice_channels", d_cur, d_cur.stride(0), zero, zero.stride(0), ctx.r, n, hw, F32)
        d_prev = torch.zeros_like(dout)
        call("uncl_splice_channels", d_prev, d_prev.stride(0), dout, dout.stride(0), ctx.r, n, hw, F32)
        return d_cur, d_prev, None


class MaxPool2(Function):
    """nn.MaxPool2d(2).  unet_parts.py:210-213."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        n, cb, h, w, _ = x.shape
        y = _empty((n, cb, h // 2, w // 2, 8), x)
        call("uncl_maxpool2", x, x.stride(0), None, 0, 0, y, y.stride(0), n, cb * 8, h, w, F32)
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        n, cb, h, w, _ = x.shape
        dx = _empty(x.shape, x)
        call("uncl_maxpool2_bwd", x, x.stride(0), dy.contiguous(), dx, n, cb * 8, h, w)
        return dx


class ConvT2x2(Function):
    """up.up: nn.ConvTranspose2d(C, C, 2, stride=2) + replicate pad to the skip size.  unet_parts.py:283-299.
    tc=True: forward and data gradient are tensor-core GEMMs on bf16-rounded operands (fp32 accumulation / outputs)."""

    @staticmethod
    def forward(ctx, x, weight, bias, h2, w2, tc=False):
        x = x.contiguous()
        n, cb, h, w, _ = x.shape
        c = cb * 8
        y = _empty((n, cb, h2, w2, 8), x)
        b = bias.detach().float().contiguous()
        if tc:
            xb = _to_bf16(x)
            call("uncl_convT2x2_tc", xb, xb.stride(0), packing.convT2x2_tc(weight.detach()), b, y, y.stride(0), F32, n, c,
                 h, w, h2, w2)
        else:
            call("uncl_convT2x2", x, x.stride(0), None, 0, 0, packing.convT2x2(weight.detach()), b, y, y.stride(0), n, c,
                 h, w, h2, w2, F32)
        ctx.save_for_backward(x, weight)
        ctx.cfg = (h2, w2, tc)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        h2, w2, tc = ctx.cfg
        n, cb, h, w, _ = x.shape
        c = cb * 8
        s2d = _empty((n, 4 * cb, h, w, 8), x)
        s2d_b = torch.empty(s2d.shape, device=x.device, dtype=torch.bfloat16) if tc else None
        call("uncl_convT2x2_s2d", dy.contiguous(), s2d, s2d_b, n, c, h, w, h2, w2)
        dx = _empty(x.shape, x)
        # columns of the GEMM are j = pos*C + co:  dx[p, ci] = sum_j s2d[p, j] * W[ci, co, dy, dx]
        if tc:
            wt = packing.pointwise_tc(weight.detach().permute(0, 2, 3, 1).reshape(c, 4 * c, 1, 1))
            call("uncl_pw_conv_tc", s2d_b, s2d_b.stride(0), wt, None, None, 0, F32, None, dx, dx.stride(0), F32, n, 4 * c,
                 c, 1, h, w, ACT_NONE)
        else:
            wt = weight.detach().permute(2, 3, 1, 0).reshape(1, 4 * c, c).contiguous().float()   # [1][4C][C]
            call("uncl_pw_conv", s2d, wt, None, None, None, dx, dx.stride(0), n, 4 * c, c, 1, h * w, ACT_NONE, F32)
        dwp = _zeros((1, c, 4 * c), x)
        call("uncl_pw_wgrad", x, s2d, dwp, n, c, 4 * c, 1, h * w)
        dw = dwp.reshape(c, 2, 2, c).permute(0, 3, 1, 2).contiguous()
        db4 = _zeros(4 * c, x)
        call("uncl_relu_bwd_bias", s2d, None, 0, db4, n, 4 * c, h * w, 0)
        return dx, dw, db4.reshape(4, c).sum(0), None, None, None


class SkipConcat(Function):
    """cat([x2, x1, x2^2, sqrt(x2 + 1e-8)], dim=1).  unet_parts.py:319-322."""

    @staticmethod
    def forward(ctx, x2, x1):
        x2, x1 = x2.contiguous(), x1.contiguous()
        n, cb, h, w, _ = x2.shape
        cat = _empty((n, 4 * cb, h, w, 8), x2)
        call("uncl_skip_concat_fwd", x2, x2.stride(0), x1, cat, n, cb * 8, h * w)
        ctx.save_for_backward(x2)
        return cat

    @staticmethod
    def backward(ctx, dcat):
        (x2,) = ctx.saved_tensors
        n, cb, h, w, _ = x2.shape
        dx2, dx1 = _empty(x2.shape, x2), _empty(x2.shape, x2)
        call("uncl_skip_concat_bwd", dcat.contiguous(), x2, x2.stride(0), dx2, dx1, n, cb * 8, h * w)
        return dx2, dx1


class AddPos(Function):
    """x + pos_embed (Unet_singleFrame.py:94); output is the [N, C/8, 144, 8] tensor of the graph block."""

    @staticmethod
    def forward(ctx, x, pos_embed):
        x = x.contiguous()
        n, cb = x.shape[0], x.shape[1]
        out = _empty((n, cb, 144, 8), x)
        call("uncl_gcn_add_pos", x, x.stride(0), packing.blocked_param(pos_embed.detach()), out, n, cb * 8, F32)
        ctx.shape = x.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = dout.contiguous()
        n, cb = dout.shape[0], dout.shape[1]
        dpos_b = _empty((cb, 144, 8), dout)
        call("uncl_batch_sum", dout, dpos_b, n, cb * 144 * 8)
        dpos = dpos_b.permute(0, 2, 1).reshape(1, cb * 8, 12, 12)
        return dout.reshape(ctx.shape), dpos


class PwConv(Function):
    """1x1 conv (+groups) with optional GELU, residual and per-sample DropPath scale:
    out = scale[n] * act(W x + b) + res.  gcn_lib/torch_vertex.py:219-227, torch_nn.py:54-78, Unet_singleFrame.py:36-42.
    tc=True: forward, data gradient and weight gradient are tensor-core GEMMs on bf16-rounded operands (fp32 accumulation /
    outputs).  tc="split": the forward is the three-term bf16 split GEMM x_hi.w_hi + x_hi.w_lo + x_lo.w_hi (~2^-16, for fc1,
    whose output drives the discrete KNN selection); gradients as tc=True."""

    @staticmethod
    def forward(ctx, x, weight, bias, res, scale, groups, gelu, tc=False):
        x = x.contiguous()
        n, cbi, hw, _ = x.shape
        ci, co = cbi * 8, weight.shape[0]
        b = bias.detach().float().contiguous()
        out = _empty((n, co // 8, hw, 8), x)
        u = None
        if gelu:
            assert res is None and scale is None
            u = _empty(out.shape, x)
        r = res.contiguous() if res is not None else None
        xb = None
        if tc:
            assert hw == 144
            xb = _to_bf16(x)
            if tc == "split":
                assert groups == 1
                xs = torch.empty((n, 3 * cbi, hw, 8), device=x.device, dtype=torch.bfloat16)
                call("uncl_split_bf16", x, x.stride(0), xs, xs.stride(0), n, ci, hw)
                call("uncl_pw_conv_tc", xs, xs.stride(0), packing.pointwise_tc_split(weight.detach()), b, r,
                     r.stride(0) if r is not None else 0, F32, scale, u if gelu else out, out.stride(0), F32, n, 3 * ci, co, 1,
                     12, 12, ACT_NONE)
            else:
                wq = packing.pointwise_tc(weight.detach(), groups)
                call("uncl_pw_conv_tc", xb, xb.stride(0), wq, b, r, r.stride(0) if r is not None else 0, F32, scale,
                     u if gelu else out, out.stride(0), F32, n, ci, co, groups, 12, 12, ACT_NONE)
        else:
            wq = packing.pointwise(weight.detach(), groups)
            call("uncl_pw_conv", x, wq, b, r, scale, u if gelu else out, out.stride(0), n, ci, co, groups, hw, ACT_NONE, F32)
        if gelu:
            call("uncl_gelu_fwd", u, out, u.numel())
        ctx.save_for_backward(xb if tc else x, weight, u, scale)
        ctx.cfg = (groups, gelu, res is not None, bool(tc))
        return out

    @staticmethod
    def backward(ctx, dout):
        x, weight, u, scale = ctx.saved_tensors
        groups, gelu, has_res, tc = ctx.cfg
        n, cbi, hw, _ = x.shape
        ci, co = cbi * 8, weight.shape[0]
        dout = dout.contiguous()
        d = dout
        if scale is not None:
            d = dout.clone()
            call("uncl_scale_rows", d, scale, n, co * hw)
        if gelu:
            du = _empty(d.shape, d)
            call("uncl_gelu_bwd", u, d, du, d.numel())
            d = du
        dx = None
        db16 = _to_bf16(d) if tc else None
        if ctx.needs_input_grad[0]:
            dx = _empty(x.shape, d)
            if tc:
                # transposed 1x1 conv: per group [C_out/g -> C_in/g], i.e. a conv with weight [C_in][C_out/g]
                wt = weight.detach().reshape(groups, co // groups, ci // groups).transpose(1, 2).reshape(ci, co // groups, 1, 1)
                call("uncl_pw_conv_tc", db16, db16.stride(0), packing.pointwise_tc(wt, groups), None, None, 0, F32, None, dx,
                     dx.stride(0), F32, n, co, ci, groups, 12, 12, ACT_NONE)
            else:
                wt = packing.pointwise(weight.detach(), groups).transpose(1, 2).contiguous()   # [g][Cout_g][Cin_g]
                call("uncl_pw_conv", d, wt, None, None, None, dx, dx.stride(0), n, co, ci, groups, hw, ACT_NONE, F32)
        dwp = _zeros((groups, ci // groups, co // groups), d)
        if tc:
            # weight gradient on the tensor cores, one GEMM over pixels per group (channel-block slices of x and dz)
            cig, cog = ci // groups, co // groups
            for g in range(groups):
                xg, zg = x[:, g * cig // 8:(g + 1) * cig // 8], db16[:, g * cog // 8:(g + 1) * cog // 8]
                call("uncl_pw_wgrad_tc", xg, x.stride(0), zg, db16.stride(0), dwp[g], n, cig, cog, 12, 12)
        else:
            call("uncl_pw_wgrad", x, d, dwp, n, ci, co, groups, hw)
        dw = dwp.permute(0, 2, 1).reshape(co, ci // groups, 1, 1).contiguous()
        db = _zeros(co, d)
        call("uncl_relu_bwd_bias", d, None, 0, db, n, co, hw, 0)
        return dx, dw, db, (dout if has_res else None), None, None, None, None


class KnnAggregate(Function):
    """KNN graph (no gradient, as in the reference: torch_edge.py:54-86 runs under no_grad) + MRConv aggregation."""

    @staticmethod
    def forward(ctx, y, relpos):
        y = y.contiguous()
        n, cb = y.shape[0], y.shape[1]
        z = _empty((n, 2 * cb, 144, 8), y)
        idx = torch.empty((n, 144, 9), device=y.device, dtype=torch.int32)
        call("uncl_gcn_knn_aggregate", y, relpos, z, F32, idx, n, cb * 8)
        ctx.save_for_backward(y, idx)
        return z

    @staticmethod
    def backward(ctx, dz):
        y, idx = ctx.saved_tensors
        n, cb = y.shape[0], y.shape[1]
        dy = _zeros(y.shape, y)
        call("uncl_gcn_agg_bwd", dz.contiguous(), y, idx, dy, n, cb * 8)
        return dy, None


class OutcSigmoid(Function):
    """outconv 1x1 (C -> 1) + Sigmoid: blocked features -> [N,1,H,W].  unet_parts.py:338-345, Unet_singleFrame.py:207-209."""

    @staticmethod
    def forward(ctx, up, weight, bias):
        up = up.contiguous()
        n, cb, h, w, _ = up.shape
        out = _empty((n, 1, h, w), up)
        wv = weight.detach().reshape(-1).float().contiguous()
        call("uncl_outc_sigmoid", up, up.stride(0), wv, bias.detach().float().contiguous(), out, None, n, cb * 8, h * w, F32)
        ctx.save_for_backward(up, out, wv)
        return out

    @staticmethod
    def backward(ctx, dout):
        up, out, wv = ctx.saved_tensors
        n, cb, h, w, _ = up.shape
        c = cb * 8
        dup = _empty(up.shape, up)
        dw, db = _zeros(c, up), _zeros(1, up)
        call("uncl_outc_sigmoid_bwd", dout.contiguous().float(), out, up, up.stride(0), wv, dup, dw, db, n, c, h * w)
        return dup, dw.reshape(1, c, 1, 1), db


class BlockedToNCHW(Function):
    """C8-blocked (fp32 or bf16) -> NCHW fp32 at the module boundary; the gradient returns in the input's dtype."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        n, cb, h, w, _ = x.shape
        o = _empty((n, cb * 8, h, w), x)
        call("uncl_blocked_to_nchw", x, x.stride(0), o, n, cb * 8, h * w, _lib.DTYPE_OF[x.dtype])
        ctx.dtype = x.dtype
        return o

    @staticmethod
    def backward(ctx, do):
        do = do.contiguous().float()
        n, c, h, w = do.shape
        dx = torch.empty((n, c // 8, h, w, 8), device=do.device, dtype=ctx.dtype)
        call("uncl_nchw_to_blocked", do, dx, dx.stride(0), n, c, h * w, _lib.DTYPE_OF[ctx.dtype])
        return dx
