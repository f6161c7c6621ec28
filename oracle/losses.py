"""Oracle: training losses.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference: models/struct_loss.py:46-104 (pyramid structural loss), GanTrainerImg.py:219-229
(contrastive_D_loss), :410-439 (nce), :308-313 (mean / contrast L1 terms), GanTrainer.py:669-682 (L_TV).
"""
import torch
import torch.nn.functional as F

from .discriminator import contrast_map

EPS2 = 1e-5  # utils/params.py:53 (params.epsilon2)


def _struct_level(a, b, ws=5):
    # models/struct_loss.py:57-87: z-normalised 5x5 windows of both images, MSE between them
    def norm_windows(img):
        window = torch.ones((1, 1, ws, ws), dtype=img.dtype) / (ws * ws)
        mu = F.conv2d(img, window)
        var = F.conv2d(img * img, window) - mu * mu
        std = torch.sqrt(torch.clamp(var, min=0) + EPS2)
        win = img.unfold(2, ws, 1).unfold(3, ws, 1)  # [N,1,H-4,W-4,5,5]
        return (win - mu[..., None, None]) / (std[..., None, None] + EPS2)

    return F.mse_loss(norm_windows(a), norm_windows(b))


def struct_loss(fake, hdr_input, pyramid_weights=(1.0, 1.0, 1.0), ws=5):
    """StructLoss.forward(fake, _, hdr_input, weights): struct_loss.py:23-54 (crop with 0 addition is a no-op)."""
    total = 0.0
    for w in pyramid_weights:
        total = total + w * _struct_level(fake, hdr_input, ws)
        fake = F.interpolate(fake, scale_factor=0.5, mode="bicubic", align_corners=False)
        hdr_input = F.interpolate(hdr_input, scale_factor=0.5, mode="bicubic", align_corners=False)
    return total


def contrastive_d_loss(real_logits, fake_logits):
    """GanTrainerImg.py:219-229: CE([r_i, f_0..f_{B-1}], 0) + CE([-f_i, -r_0..-r_{B-1}], 0)."""
    r = real_logits.reshape(-1)
    f = fake_logits.reshape(-1)

    def half(t1, t2):
        t = torch.cat([t1[:, None], t2[None, :].expand(t1.shape[0], -1)], dim=1)
        return F.cross_entropy(t, torch.zeros(t1.shape[0], dtype=torch.long))

    return half(r, f) + half(-f, -r)


def nce(anchor, positive, negative, k, constant):
    """GanTrainerImg.py:410-439 with one positive and one negative, cl_loss_type='InfoNCE'."""
    def sim(a, b):
        return torch.sum((a * b) * (1 / (constant + k * torch.abs(a - b))), dim=1).mean(dim=(-1, -2)).unsqueeze(1)

    logits = torch.cat([sim(anchor, positive), sim(anchor, negative)], dim=1)
    return F.cross_entropy(logits, torch.zeros(anchor.shape[0], dtype=torch.long))


def l1_mean_terms(fake, ldr):
    """GanTrainerImg.py:308-313: L1 of per-image means and of per-image mean local variance."""
    l_mean = F.l1_loss(fake.mean(dim=(-1, -2)), ldr.mean(dim=(-1, -2)))
    l_con = F.l1_loss(contrast_map(fake).mean(dim=(-1, -2)), contrast_map(ldr).mean(dim=(-1, -2)))
    return l_mean, l_con


def tv_loss(x):
    """GanTrainer.py:669-682 (L_TV, weight 1)."""
    b, _, h, w = x.shape
    h_tv = ((x[:, :, 1:, :] - x[:, :, :-1, :]) ** 2).sum()
    w_tv = ((x[:, :, :, 1:] - x[:, :, :, :-1]) ** 2).sum()
    return 2 * (h_tv / ((h - 1) * w) + w_tv / (h * (w - 1))) / b
