"""Deterministic reference-shaped parameters for the generator and discriminator.

The reference ships no trained weights, so every parity run uses random-init
weights of the shipped architecture.  This module produces a `state_dict` with the exact keys and
shapes of SURVEY.md Appendix B and the same init *distributions* the reference applies
(`weights_init_xavier` on Conv2d/Linear, utils/model_save_util.py:41-47; PyTorch default init on
ConvTranspose2d), from an explicit seed, so the same tensors can be loaded into the reference modules
(golden generation), the oracle and the B200 modules.
"""
import math

import numpy as np
import torch

FILTERS = 32
GRID = 12  # bottleneck is 12x12 (Unet_singleFrame.py:66)


def generator_shapes():
    """(key, shape, kind) for every generator parameter, in state_dict order."""
    f = FILTERS
    out = [("inc.conv.conv", (f, 1, 3, 3), "conv"), ("inc.conv.conv1", (f, f, 3, 3), "conv")]
    ch = f
    for i in range(3):
        out += [("down_path.%d.mpconv.1.conv" % i, (2 * ch, ch, 3, 3), "conv"),
                ("down_path.%d.mpconv.1.conv1" % i, (2 * ch, 2 * ch, 3, 3), "conv")]
        ch *= 2
    out += [("down_path.3.mpconv.1.conv", (ch, ch, 3, 3), "conv"),
            ("down_path.3.mpconv.1.conv1", (ch, ch, 3, 3), "convT")]
    g = "gcn.module.0."
    out += [(g + "0.fc1.0", (ch, ch, 1, 1), "conv"),
            (g + "0.graph_conv.gconv.nn.0", (2 * ch, 2 * ch // 4, 1, 1), "conv"),
            (g + "0.fc2.0", (ch, 2 * ch, 1, 1), "conv"),
            (g + "1.fc1.0", (ch, ch, 1, 1), "conv"),
            (g + "1.fc2.0", (ch, ch, 1, 1), "conv")]
    for i in range(4):
        co = f if i >= 2 else ch // 2
        out += [("up_path.%d.up" % i, (ch, ch, 2, 2), "convT"),
                ("up_path.%d.conv.conv" % i, (4 * ch, co, 3, 3), "convT"),
                ("up_path.%d.conv.conv1" % i, (co, co, 3, 3), "convT")]
        ch //= 2
    out += [("outc.conv", (1, f, 1, 1), "conv")]
    return out


def relative_pos_table(channels=256, grid=GRID):
    """Fixed KNN bias of the Grapher: -(2 E E^T / D), E = 2-D sin/cos position embedding.

    gcn_lib/torch_vertex.py:203-209, gcn_lib/pos_embed.py:21-85 (the bicubic resize to (n, n) is the identity).
    """
    half = channels // 2
    omega = 1.0 / 10000 ** (np.arange(half // 2, dtype=np.float64) / (half / 2.0))
    gw, gh = np.meshgrid(np.arange(grid, dtype=np.float32), np.arange(grid, dtype=np.float32))

    def emb(pos):
        o = pos.reshape(-1).astype(np.float64)[:, None] * omega[None, :]
        return np.concatenate([np.sin(o), np.cos(o)], axis=1)

    e = np.concatenate([emb(gw), emb(gh)], axis=1)
    return -torch.from_numpy(np.float32(2.0 * (e @ e.T) / e.shape[1])).unsqueeze(0)


def _fill(shape, kind, gen, bias_scale):
    if kind == "conv":  # xavier_normal_(gain=sqrt(2)); bias 0 in the reference, small noise here if asked
        rf = shape[2] * shape[3]
        std = math.sqrt(2.0) * math.sqrt(2.0 / (shape[1] * rf + shape[0] * rf))
        w = torch.randn(shape, generator=gen) * std
        b = torch.randn(shape[0], generator=gen) * (bias_scale * std)
    else:  # ConvTranspose2d default: kaiming_uniform(a=sqrt(5)) -> U(-1/sqrt(fan_in), ..), fan_in = shape[1]*k*k
        bound = 1.0 / math.sqrt(shape[1] * shape[2] * shape[3])
        w = (torch.rand(shape, generator=gen) * 2 - 1) * bound
        b = (torch.rand(shape[1], generator=gen) * 2 - 1) * bound
    return w, b


def make_generator_state_dict(seed=999, bias_scale=1.0, pos_embed_std=0.02):
    """Reference-shaped generator parameters.  bias_scale / pos_embed_std > 0 perturb the tensors the
    reference zero-initialises so that parity tests exercise every term; pass 0 for the literal init."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape, kind in generator_shapes():
        w, b = _fill(shape, kind, gen, bias_scale)
        sd[key + ".weight"] = w
        sd[key + ".bias"] = b
        if key.endswith("mpconv.1.conv1") and key.startswith("down_path.3"):
            sd["gcn.pos_embed"] = torch.randn(1, 8 * FILTERS, GRID, GRID, generator=gen) * pos_embed_std
            sd["gcn.module.0.0.relative_pos"] = relative_pos_table(8 * FILTERS, GRID)
    return sd


def make_discriminator_state_dict(seed=1999, bias_scale=1.0):
    """SimpleDiscriminator parameters (models/Discriminator.py:87-126, dim=16, input 256)."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape in (("model.0", (16, 1, 4, 4)), ("model.2", (32, 16, 4, 4)), ("model.4", (1, 32, 1, 1))):
        w, b = _fill(shape, "conv", gen, bias_scale)
        sd[key + ".weight"], sd[key + ".bias"] = w, b
    sd["tail.1.weight"] = torch.randn(1, 62 * 62, generator=gen) * math.sqrt(2.0) * math.sqrt(2.0 / (62 * 62 + 1))
    return sd
