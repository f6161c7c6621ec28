"""Oracle: SimpleDiscriminator forward.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference: models/Discriminator.py:87-126 (SimpleDiscriminator), :61-83 (compute_contrast /
ContrastExtracter), :49-55 (fspecial_gauss); shipped config d_down_dim=16, d_padding=0,
simpleD_maxpool=0, d_last_activation='none', d_norm='none'.
"""
import numpy as np
import torch
import torch.nn.functional as F


def gauss_window(size=11, sigma=1.5):
    ax = np.arange(-(size // 2), size // 2 + 1)
    xx, yy = np.meshgrid(ax, ax, indexing="ij")
    g = np.exp(-((xx ** 2 + yy ** 2) / (2.0 * sigma ** 2)))
    return torch.from_numpy(g / g.sum()).float().unsqueeze(0).unsqueeze(0)


def contrast_map(x, win=None):
    """Local variance under an 11x11 gaussian (valid): E[x^2] - E[x]^2, per channel."""
    b, c, h, w = x.shape
    win = (gauss_window() if win is None else win).to(x)   # dtype and device of x
    xr = x.reshape(b * c, 1, h, w)
    mu = F.conv2d(xr, win)
    var = F.conv2d(xr * xr, win) - mu * mu
    return var.reshape(b, c, var.shape[2], var.shape[3])


def simple_discriminator_forward(sd, x):
    """x [N,1,256,256] -> (logit [N,1], fea [N,2,1,1]);  Discriminator.py:119-126."""
    h = F.leaky_relu(F.conv2d(x, sd["model.0.weight"], sd["model.0.bias"], stride=2), 0.2)
    h = F.leaky_relu(F.conv2d(h, sd["model.2.weight"], sd["model.2.bias"], stride=2), 0.2)
    fea = F.conv2d(h, sd["model.4.weight"], sd["model.4.bias"])  # [N,1,62,62]
    out = F.linear(fea.flatten(1), sd["tail.1.weight"])
    fea1 = fea.mean(dim=(2, 3), keepdim=True)
    fea2 = contrast_map(fea).mean(dim=(2, 3), keepdim=True)
    return out, torch.cat([fea1, fea2], dim=1)
