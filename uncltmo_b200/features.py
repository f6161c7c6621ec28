"""Mean / local-contrast features shared by the video generator, the discriminator and the L1 loss terms."""
import torch

from ._lib import call


def plane_mean_contrast(x, want_mean=True, want_contrast=True):
    """x [..., H, W] fp32 CUDA (dense planes) -> (mean [...], mean local variance [...]).

    ContrastExtracter + adaptive_avg_pool2d: models/Discriminator.py:61-83,121-125; Unet.py:112-133,274-278."""
    from .autograd_losses import PlaneMeanContrastFn
    return PlaneMeanContrastFn.apply(x)


def contrast_features(up_blocked):
    """Video-generator features: cat[avgpool(up_x), avgpool(local variance(up_x))] -> [N, 2C, 1, 1] (Unet.py:274-278)."""
    from .generator import blocked_to_nchw
    x = blocked_to_nchw(up_blocked)
    mean, con = plane_mean_contrast(x)
    return torch.cat([mean, con], dim=1)[:, :, None, None]
