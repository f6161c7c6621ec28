"""Oracle: unet_multi_filters generator (image + video), functional restatement.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  All functions take the reference
`state_dict` (keys of SURVEY.md Appendix B) and NCHW tensors on CPU.

Reference: models/unet_multi_filters/Unet_singleFrame.py:177-213 (image forward),
Unet.py:213-289 (video forward), unet_parts.py (blocks), gcn_lib/* (Grapher).
"""
import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-8  # utils/params.py:52 (params.epsilon)

# ---------------------------------------------------------------------------------------------------------------
# bf16-operand emulation.  NOT reference behaviour (the reference is fp32 throughout): with `bf16_operands(True)` the
# oracle rounds, to bfloat16 and back, exactly the tensors the mixed-precision CUDA path stores in bf16 - the
# activation and weight operands of every tensor-core convolution and the pre-activation gradients dz that feed the
# data- and weight-gradient GEMMs - while every accumulation stays in the oracle's dtype (float64 in the tests).  The
# remaining difference to the CUDA path is fp32 accumulation order, so gradient parity can be gated per tensor instead
# of by cosine similarity.
_BF16_OPERANDS = False


class _Round(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


class _GradRound(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


class bf16_operands:
    """Context manager: `with oracle.bf16_operands(True): ...` (see above)."""

    def __init__(self, on=True):
        self.on = bool(on)

    def __enter__(self):
        global _BF16_OPERANDS
        self.prev, _BF16_OPERANDS = _BF16_OPERANDS, self.on

    def __exit__(self, *a):
        global _BF16_OPERANDS
        _BF16_OPERANDS = self.prev


def _r(t):
    """operand stored / read as bf16 by the tensor-core path"""
    return _Round.apply(t) if _BF16_OPERANDS else t


def _conv(x, w, b, transposed=False, **kw):
    """A convolution that the CUDA path runs on the tensor cores: bf16 operands, pre-activation gradient in bf16."""
    if not _BF16_OPERANDS:
        return (F.conv_transpose2d if transposed else F.conv2d)(x, w, b, **kw)
    z = (F.conv_transpose2d if transposed else F.conv2d)(_Round.apply(x), _Round.apply(w), None, **kw)
    z = _GradRound.apply(z)
    return z if b is None else z + b.view(1, -1, 1, 1)


def _double_conv(sd, p, x):
    # unet_parts.py:57-87 with padding=0, unet_norm='none', activation='relu'
    first = sd[p + "conv.weight"].shape[1] == 1   # inc.conv.conv (1 -> 32): fp32 CUDA cores in both precision modes
    x = F.relu((F.conv2d if first else _conv)(x, sd[p + "conv.weight"], sd[p + "conv.bias"]))
    return F.relu(_conv(x, sd[p + "conv1.weight"], sd[p + "conv1.bias"]))


def _double_last_conv(sd, p, x):
    # unet_parts.py:126-141: conv3 valid -> ReLU -> ConvTranspose 3x3 s1 p0 -> ReLU
    x = F.relu(_conv(x, sd[p + "conv.weight"], sd[p + "conv.bias"]))
    return F.relu(_conv(x, sd[p + "conv1.weight"], sd[p + "conv1.bias"], transposed=True))


def _up(sd, p, x1, x2):
    # unet_parts.py:283-335, up_mode=0, convtranspose_kernel=2, con_operator='square_and_square_root'
    x1 = _conv(x1, sd[p + "up.weight"], sd[p + "up.bias"], transposed=True, stride=2)
    dy = x2.shape[2] - x1.shape[2]
    dx = x2.shape[3] - x1.shape[3]
    if dx or dy:
        x1 = F.pad(x1, (dx // 2, dx - dx // 2, dy // 2, dy - dy // 2), mode="replicate")
    x = torch.cat([x2, x1, x2 * x2, torch.pow(x2 + EPS, 0.5)], dim=1)  # :319-322
    # double_conv_traspose, unet_parts.py:183-193
    x = F.relu(_conv(x, sd[p + "conv.conv.weight"], sd[p + "conv.conv.bias"], transposed=True))
    return F.relu(_conv(x, sd[p + "conv.conv1.weight"], sd[p + "conv.conv1.bias"], transposed=True))


def relative_pos_table(channels=256, grid=12):
    """The fixed `relative_pos` buffer: gcn_lib/torch_vertex.py:203-209 + pos_embed.py:21-85.

    -(2 * E E^T / D) with E the 2-D sin/cos embedding (w first); the bicubic resize to (n, n)
    is the identity at r=1.
    """
    half = channels // 2
    omega = 1.0 / 10000 ** (np.arange(half // 2, dtype=np.float64) / (half / 2.0))
    gw, gh = np.meshgrid(np.arange(grid, dtype=np.float32), np.arange(grid, dtype=np.float32))

    def emb1d(pos):
        out = np.einsum("m,d->md", pos.reshape(-1).astype(np.float64), omega)
        return np.concatenate([np.sin(out), np.cos(out)], axis=1)

    e = np.concatenate([emb1d(gw), emb1d(gh)], axis=1)  # (n, D); "w goes first" (pos_embed.py:44)
    rel = 2.0 * (e @ e.T) / e.shape[1]
    return -torch.from_numpy(np.float32(rel)).unsqueeze(0)


def knn_indices(y, relative_pos, k=9):
    """DenseDilatedKnnGraph (dilation 1): gcn_lib/torch_edge.py:135-159, 54-86, 9-20.

    y: [B, C, n] un-normalised node features.  Returns idx [B, n, k] (int64).
    """
    yn = F.normalize(y, p=2.0, dim=1).transpose(2, 1)  # [B, n, C]
    inner = -2 * torch.matmul(yn, yn.transpose(2, 1))
    sq = torch.sum(yn * yn, dim=-1, keepdim=True)
    dist = sq + inner + sq.transpose(2, 1)
    if relative_pos is not None:
        dist = dist + relative_pos
    return torch.topk(-dist, k=k)[1]


def gcn_block(sd, x, droppath_masks=None, return_idx=False):
    """GCNBlock: Unet_singleFrame.py:93-99; Grapher_noBN torch_vertex.py:219-227; FFN :36-42.

    droppath_masks: None (eval / p=0) or a pair of [B] tensors of per-sample scale factors
    (mask/keep_prob) applied to the two residual branches (train-mode DropPath, p=0.05).
    """
    p = "gcn.module.0."
    x = x + sd["gcn.pos_embed"]
    B, C, H, W = x.shape
    n = H * W
    y = F.conv2d(x, sd[p + "0.fc1.0.weight"], sd[p + "0.fc1.0.bias"]).reshape(B, C, n)
    idx = knn_indices(y.detach(), sd[p + "0.relative_pos"])  # [B, n, 9]
    # MRConv2d, torch_vertex.py:21-30: max_k (y_j - y_i), interleaved with y channel-wise
    yj = torch.gather(y.unsqueeze(-1).expand(B, C, n, idx.shape[-1]), 2,
                      idx.unsqueeze(1).expand(B, C, n, idx.shape[-1]))
    agg = (yj - y.unsqueeze(-1)).max(dim=-1)[0]
    z = torch.stack([y, agg], dim=2).reshape(B, 2 * C, H, W)
    z = F.gelu(_conv(z, sd[p + "0.graph_conv.gconv.nn.0.weight"],
                     sd[p + "0.graph_conv.gconv.nn.0.bias"], groups=4))
    z = _conv(z, sd[p + "0.fc2.0.weight"], sd[p + "0.fc2.0.bias"])
    if droppath_masks is not None:
        z = z * droppath_masks[0].view(B, 1, 1, 1)
    x = z + x
    f = F.gelu(_conv(x, sd[p + "1.fc1.0.weight"], sd[p + "1.fc1.0.bias"]))
    f = _conv(f, sd[p + "1.fc2.0.weight"], sd[p + "1.fc2.0.bias"])
    if droppath_masks is not None:
        f = f * droppath_masks[1].view(B, 1, 1, 1)
    out = f + x
    return (out, idx) if return_idx else out


def _encode(sd, x_in, prev=None, cur=None):
    """inc + 4 down stages.  prev/cur: recurrent slices of the video generator (Unet.py:229-251)."""
    nxt = _double_conv(sd, "inc.conv.", x_in)
    skips = [nxt]
    if cur is not None:
        cur.append(nxt[:, :nxt.shape[1] // 32])
    for i in range(4):
        fea = nxt
        if prev is not None:
            r = nxt.shape[1] // 32
            fea = torch.cat((prev[i], nxt[:, r:]), 1)  # Unet.py:244
        fea = F.max_pool2d(fea, 2)
        pre = "down_path.%d.mpconv.1." % i
        nxt = _double_conv(sd, pre, fea) if i < 3 else _double_last_conv(sd, pre, fea)
        skips.append(nxt)
        if cur is not None:
            cur.append(nxt[:, :nxt.shape[1] // 32])
    return skips


def unet_forward(sd, x, droppath_masks=None, return_all=False):
    """Image generator forward, Unet_singleFrame.py:177-213 (shipped hyper-parameters).

    x: [N,1,256,256] -> (sigmoid map [N,1,256,256], up_x [N,32,256,256]).
    """
    skips = _encode(sd, x)
    up_x = gcn_block(sd, skips[4], droppath_masks)
    inter = {"skips": skips, "gcn": up_x, "ups": []}
    for i in range(4):
        up_x = _up(sd, "up_path.%d." % i, up_x, skips[3 - i])
        inter["ups"].append(up_x)
    up_x = _r(up_x)   # stored as bf16 by the tensor-core path; the out conv and the feature consumers read that tensor
    logit = F.conv2d(up_x, sd["outc.conv.weight"], sd["outc.conv.bias"])
    out = torch.sigmoid(logit)
    if return_all:
        inter["logit"] = logit
        return out, up_x, inter
    return out, up_x


def gauss_window11():
    # Unet.py:101-106 fspecial_gauss(11, 1.5)
    ax = np.arange(-5, 6)
    xx, yy = np.meshgrid(ax, ax, indexing="ij")
    g = np.exp(-((xx ** 2 + yy ** 2) / (2.0 * 1.5 ** 2)))
    return torch.from_numpy(g / g.sum()).float().unsqueeze(0).unsqueeze(0)


def _contrast_features(up_x):
    # Unet.py:274-278, compute_contrast :112-123
    b, c, h, w = up_x.shape
    win = gauss_window11().to(up_x.dtype)
    xr = up_x.reshape(b * c, 1, h, w)
    mu = F.conv2d(xr, win)
    var = F.conv2d(xr * xr, win) - mu * mu
    var = var.reshape(b, c, var.shape[2], var.shape[3])
    return torch.cat([up_x.mean(dim=(2, 3), keepdim=True), var.mean(dim=(2, 3), keepdim=True)], dim=1)


def unet_video_forward(sd, x, droppath_masks=None, detach_state=False):
    """Video generator forward, Unet.py:213-289.

    x: [N,T,1,256,256] -> (frames [N,T,1,256,256], features [N,T,64,1,1]).  Frame k feeds the first
    C/32 channels of its 8 stage inputs from frame k-1 (not detached).
    droppath_masks: optional list (per frame) of mask pairs.
    detach_state: NOT the reference behaviour - cuts the gradient through the hand-over, so that tests can measure how
    much of a gradient travels through the recurrence.
    """
    outs, feats, prev = [], [], None
    for k in range(x.shape[1]):
        cur = []
        skips = _encode(sd, x[:, k], prev, cur)
        up_x = gcn_block(sd, skips[4], None if droppath_masks is None else droppath_masks[k])
        cur.append(up_x[:, :up_x.shape[1] // 32])
        for i in range(4):
            fea = up_x
            if prev is not None:
                r = up_x.shape[1] // 32
                fea = torch.cat((prev[5 + i], up_x[:, r:]), 1)  # Unet.py:270
            up_x = _up(sd, "up_path.%d." % i, fea, skips[3 - i])
            cur.append(up_x[:, :up_x.shape[1] // 32])
        up_x = _r(up_x)
        feats.append(_contrast_features(up_x).unsqueeze(1))
        out = torch.sigmoid(F.conv2d(up_x, sd["outc.conv.weight"], sd["outc.conv.bias"]))
        outs.append(out.unsqueeze(1))
        prev = [c.detach() for c in cur] if detach_state else cur
    return torch.cat(outs, 1), torch.cat(feats, 1)


GFLOP_PER_TILE = 18.2858  # counted on the reference module
