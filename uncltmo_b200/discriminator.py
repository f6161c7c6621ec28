"""Drop-in SimpleDiscriminator (models/Discriminator.py:87-126) on sm_100a kernels, forward and backward.

Same constructor signature, `forward(x) -> (output [N,1], fea_final [N,2,1,1])` and state_dict keys
(`model.0/2/4.{weight,bias}`, `tail.1.weight`).  Only the shipped configuration is built
(d_model='simpleD', dim 16, input 256, padding 0, no max-pool, last_activation 'none', norm 'none').
"""
import torch
import torch.nn as nn

from .autograd_losses import DiscFn, PlaneMeanContrastFn


class SimpleDiscriminator(nn.Module):
    def __init__(self, input_size, input_dim, dim, norm, last_activation, simpleD_maxpool, padding):
        super().__init__()
        if (input_size, input_dim, dim, norm, last_activation, bool(simpleD_maxpool), padding) != (256, 1, 16, "none", "none", False, 0):
            raise NotImplementedError("uncltmo_b200 builds the shipped SimpleDiscriminator configuration only")
        self.model = nn.ModuleList([nn.Conv2d(1, 16, 4, 2), nn.Identity(), nn.Conv2d(16, 32, 4, 2), nn.Identity(),
                                    nn.Conv2d(32, 1, 1)])
        self.tail = nn.ModuleList([nn.Identity(), nn.Linear(62 * 62, 1, bias=False)])

    def forward(self, x, want_features=True):
        """want_features=False (an addition to the reference signature): skip the [N,2,1,1] feature statistics when only
        the logits are used (the D step), second return value None."""
        if x.dim() != 4 or tuple(x.shape[1:]) != (1, 256, 256):
            raise ValueError("SimpleDiscriminator expects [N,1,256,256]")
        if not x.is_cuda:
            raise RuntimeError("uncltmo_b200 has no CPU path: move the input to a CUDA device")
        m = self.model
        logits, fea = DiscFn.apply(x, m[0].weight, m[0].bias, m[2].weight, m[2].bias, m[4].weight, m[4].bias,
                                   self.tail[1].weight)
        if not want_features:
            return logits, None
        mean, con = PlaneMeanContrastFn.apply(fea)
        return logits, torch.cat([mean, con], dim=1)[:, :, None, None]
