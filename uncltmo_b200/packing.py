"""One-time re-layout of reference-shaped parameters into the formats the kernels read.

These run on the device with torch tensor ops when weights are (re)loaded - plumbing, not the hot path.
Reference weight layouts: nn.Conv2d [C_out, C_in, kH, kW]; nn.ConvTranspose2d [C_in, C_out, kH, kW]
(SURVEY.md Appendix B).
"""
import torch


def conv3x3_taps(w, transposed):
    """-> [9][C_in][C_out] fp32 (tap = ky*3+kx of the equivalent plain correlation).

    transposed=True: ConvTranspose2d 3x3 s1 p0 == correlation over a 2-px zero-padded input with the kernel
    flipped in both axes (unet_parts.py:114, 148-159)."""
    return conv3x3_taps_layout(w, transposed).float()


def conv3x3_taps_layout(w, transposed):
    """conv3x3_taps without the cast (pure permutation: also applied to index tensors, see build_pack_maps)."""
    if transposed:
        return w.flip(2, 3).permute(2, 3, 0, 1).reshape(9, w.shape[0], w.shape[1]).contiguous()
    return w.permute(2, 3, 1, 0).reshape(9, w.shape[1], w.shape[0]).contiguous()


def conv3x3_dgrad_taps_layout(w9):
    """[9][C_in][C_out] taps of a correlation -> the taps of its data gradient (a correlation of dZ with pad 2 - p):
    reversed tap order, input / output channels swapped."""
    return w9.flip(0).transpose(1, 2).contiguous()


def conv3x3_tc_is_merged(ci, co):
    """True when uncl_conv3x3_tc routes a (C_in, C_out) layer to the kx-merged kernel.  Asked of the library itself
    (uncl_conv3x3_tc_plan is pure host arithmetic), so the packing can never drift from conv_tc.cu:use_merged()."""
    import ctypes
    from . import _lib
    plan = (ctypes.c_int * 16)()
    rc = _lib.lib().uncl_conv3x3_tc_plan(1, ci, 64, 64, co, 0, ctypes.cast(plan, ctypes.c_void_p))
    if rc != 0:
        raise RuntimeError("uncl_conv3x3_tc_plan failed (%d): %s" % (rc, _lib.lib().uncl_last_error().decode()))
    return (plan[0] & 1) == 1


def conv3x3_tc(w9):
    """[9][C_in][C_out] fp32 -> the tensor-core B operand of uncl_conv3x3_tc (bf16, K-major core matrices of 8 x 16 B).

    C_out <= 64, C_in >= 64 and C_in % 32 == 0 (conv_tc_merged.cu, the three kx taps of a filter row merged into N):
        [NS][C_in/16][3 ky][2][3*NT (kx, n)][8], NT = min(C_out, 64)
    wider layers (conv_tc.cu, one tap per MMA): [NS][C_in/16][9][2][NT][8], NT = min(C_out, 128)."""
    return conv3x3_tc_layout(w9).to(torch.bfloat16)


def conv3x3_tc_layout(w9):
    _, ci, co = w9.shape
    if conv3x3_tc_is_merged(ci, co):
        nt = min(co, 64)
        ns = co // nt
        t = w9.reshape(3, 3, ci // 16, 2, 8, ns, nt)       # ky, kx, chunk, half, k8, ns, n
        t = t.permute(5, 2, 0, 3, 1, 6, 4).contiguous()    # ns, chunk, ky, half, kx, n, k8
        return t.reshape(ns, ci // 16, 3, 2, 3 * nt, 8)
    nt = min(co, 128)
    ns = co // nt
    t = w9.reshape(9, ci // 16, 2, 8, ns, nt)          # tap, chunk, half, k8, ns, n
    return t.permute(4, 1, 0, 2, 5, 3).contiguous()    # ns, chunk, tap, half, n, k8


def conv3x3_tc_rows_plan(n, ci, h, w, pad, derive=False, sms=148, co=32):
    """plan[16] of uncl_conv3x3_tc_rows_plan (pure host arithmetic, no GPU): plan[0] == 1 when the row kernel takes the
    layer, plan[2] = trailing output columns it leaves to uncl_conv3x3_tc / uncl_conv3x3_tc_skipcat."""
    import ctypes
    from . import _lib
    plan = (ctypes.c_int * 16)()
    rc = _lib.lib().uncl_conv3x3_tc_rows_plan(n, ci, h, w, co, pad, 1 if derive else 0, sms, ctypes.cast(plan, ctypes.c_void_p))
    if rc != 0:
        raise RuntimeError("uncl_conv3x3_tc_rows_plan failed (%d): %s" % (rc, _lib.lib().uncl_last_error().decode()))
    return list(plan)


def conv3x3_tc_rows(w9):
    """[9][C_in][C_out] fp32 (C_out = 32 or 64) -> the B operand of the row kernel (conv_tc_rows.cu, the three ky taps of a
    filter column merged into N): bf16 [C_in/32][2 ksteps][3 kx][2][3*C_out (ky, co)][8]."""
    return conv3x3_tc_rows_layout(w9).to(torch.bfloat16)


def conv3x3_tc_rows_layout(w9):
    _, ci, co = w9.shape
    if co not in (32, 64) or ci % 32 != 0:
        raise ValueError("the row kernel is built for C_out in (32, 64) and C_in %% 32 == 0 (got %d -> %d)" % (ci, co))
    t = w9.reshape(3, 3, ci // 32, 2, 2, 8, co)        # ky, kx, chunk, kstep, half, k8, co
    t = t.permute(2, 3, 1, 4, 0, 6, 5).contiguous()    # chunk, kstep, kx, half, ky, co, k8
    return t.reshape(ci // 32, 2, 3, 2, 3 * co, 8)


def _hi_lo(w):
    hi = w.float().to(torch.bfloat16).float()
    return hi, w.float() - hi


def conv3x3_tc_split(w9):
    """[9][C_in][C_out] fp32 -> the B operand of the exact (three-term bf16 split) tensor-core conv: input channels
    [w_hi ; w_lo ; w_hi] (3 C_in), matching uncl_split_bf16's [x_hi | x_hi | x_lo]; packed like conv3x3_tc for (3 C_in, C_out)."""
    hi, lo = _hi_lo(w9)
    return conv3x3_tc(torch.cat([hi, lo, hi], dim=1))


def convT2x2_tc_split(w):
    """ConvTranspose2d k2 s2 weight [C_in][C_out][2][2] -> convT2x2_tc layout of [w_hi ; w_lo ; w_hi] (3 C_in rows)."""
    hi, lo = _hi_lo(w)
    return convT2x2_tc(torch.cat([hi, lo, hi], dim=0))


def convT2x2(w):
    """ConvTranspose2d k2 s2 weight [C_in][C_out][2][2] -> [C_in][4][C_out] fp32 (pos = dy*2+dx)."""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], 4, w.shape[1]).contiguous().float()


def convT2x2_gemm_layout(w):
    """ConvTranspose2d k2 s2 weight [C_in][C_out][2][2] -> [C_in][4*C_out] with column j = (dy*2+dx)*C_out + co: the
    layout of the up-convolution's weight-gradient GEMM (dW[ci][j] = sum_pix X[pix][ci] * s2d(dY)[pix][j])."""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], 4 * w.shape[1]).contiguous()


def convT2x2_dgrad_layout(w):
    """The same weight as the 1x1 conv [C_out' = C_in][C_in' = 4*C_out][1][1] that maps the space-to-depth output
    gradient back to the input gradient (uncl_pw_conv_tc_dgrad)."""
    return pointwise_tc_layout(convT2x2_gemm_layout(w).reshape(w.shape[0], 4 * w.shape[1], 1, 1), 1)


def convT2x2_tc(w):
    """ConvTranspose2d k2 s2 weight [C_in][C_out][2][2] -> bf16 [NS][C_in/16][2][NT][8] (GEMM B operand, K-major):
    column j = (dy*2+dx)*C_out + co, NT = min(4*C_out, 128)."""
    return convT2x2_tc_layout(w).to(torch.bfloat16)


def convT2x2_tc_layout(w):
    ci, co = w.shape[0], w.shape[1]
    n_total = 4 * co
    nt = min(n_total, 128)
    b = w.permute(2, 3, 1, 0).reshape(n_total, ci)             # [j][ci]
    return b.reshape(n_total // nt, nt, ci // 16, 2, 8).permute(0, 2, 3, 1, 4).contiguous()


def pointwise(w, groups=1):
    """1x1 Conv2d weight [C_out][C_in/g][1][1] -> [g][C_in/g][C_out/g] fp32."""
    co, cig = w.shape[0], w.shape[1]
    return w.reshape(groups, co // groups, cig).permute(0, 2, 1).contiguous().float()


def pointwise_tc(w, groups=1):
    """1x1 Conv2d weight [C_out][C_in/g][1][1] -> bf16 [NS][C_in/g/16][2][NT][8] (conv_tc.cu pointwise B operand, K-major):
    N split ns covers output channels ns*NT .. ns*NT+NT-1 (inside one group), NT = min(C_out/g, 128)."""
    return pointwise_tc_layout(w, groups).to(torch.bfloat16)


def pointwise_tc_layout(w, groups=1):
    co, cig = w.shape[0], w.shape[1]
    nt = min(co // groups, 128)
    return w.reshape(co // nt, nt, cig // 16, 2, 8).permute(0, 2, 3, 1, 4).contiguous()


def pointwise_tc_split(w):
    """1x1 Conv2d weight [C_out][C_in][1][1] fp32 -> the B operand of the three-term bf16 GEMM x_hi.w_hi + x_hi.w_lo +
    x_lo.w_hi (input channels [w_hi | w_lo | w_hi], matching uncl_gcn_add_pos_split's [hi | hi | lo]): bf16
    [NS][3*C_in/16][2][NT][8]."""
    w2 = w.reshape(w.shape[0], w.shape[1]).float()
    hi = w2.to(torch.bfloat16)
    lo = (w2 - hi.float()).to(torch.bfloat16)
    cat = torch.cat([hi, lo, hi], dim=1).float()
    return pointwise_tc(cat.reshape(cat.shape[0], cat.shape[1], 1, 1), 1)


def conv_first(w):
    """Conv2d(1, C_out, 3) weight [C_out][1][3][3] -> [9][C_out] fp32."""
    return conv_first_layout(w).float()


def conv_first_layout(w):
    return w.reshape(w.shape[0], 9).t().contiguous()


def conv_first_rows(w, b):
    """Conv2d(1, 32, 3) weight [32][1][3][3] + bias [32] -> the B operand of the row kernel's front mode
    (uncl_conv_first_conv3x3_tc_rows): bf16 [2 (hi, lo)][2 K halves][32 (co)][8 (tap in half)]; tap 9 holds the bias (its
    input is the constant 1), taps 10..15 are zero."""
    w10 = torch.cat([w.reshape(w.shape[0], 9).t().float(), b.reshape(1, -1).float()], dim=0)     # [10][32]
    hi = w10.to(torch.bfloat16)
    lo = (w10 - hi.float()).to(torch.bfloat16)
    t = torch.zeros((2, 16, w.shape[0]), device=w.device, dtype=torch.bfloat16)
    t[0, :10], t[1, :10] = hi, lo
    return t.reshape(2, 2, 8, w.shape[0]).permute(0, 1, 3, 2).contiguous()


def blocked_param(t):
    """[1][C][H][W] -> C8-blocked [C/8][H*W][8] fp32 (pos_embed)."""
    return blocked_param_layout(t).float()


def blocked_param_layout(t):
    _, c, h, w = t.shape
    return t.reshape(c // 8, 8, h * w).permute(0, 2, 1).contiguous()


LO_FLAG = 1 << 30   # index-map flag: store the bf16 residual w - bf16(w) (uncl_pack_gather)


def pointwise_tc_split_index_layout(idx):
    """Index-map form of pointwise_tc_split: [C_out][C_in][1][1] int64 source indices -> [hi | lo | hi] B operand."""
    i2 = idx.reshape(idx.shape[0], idx.shape[1])
    cat = torch.cat([i2, i2 + LO_FLAG, i2], dim=1)
    return pointwise_tc_layout(cat.reshape(cat.shape[0], cat.shape[1], 1, 1), 1)
