"""Drop-in SimpleDiscriminator (models/Discriminator.py:87-126), forward on sm_100a kernels.

Same constructor signature, `forward(x) -> (output [N,1], fea_final [N,2,1,1])` and state_dict keys
(`model.0/2/4.{weight,bias}`, `tail.1.weight`).  Only the shipped configuration is built
(d_model='simpleD', dim 16, input 256, padding 0, no max-pool, last_activation 'none', norm 'none').
"""
import torch
import torch.nn as nn

from ._lib import call
from .features import plane_mean_contrast


class SimpleDiscriminator(nn.Module):
    def __init__(self, input_size, input_dim, dim, norm, last_activation, simpleD_maxpool, padding):
        super().__init__()
        if (input_size, input_dim, dim, norm, last_activation, bool(simpleD_maxpool), padding) != (256, 1, 16, "none", "none", False, 0):
            raise NotImplementedError("uncltmo_b200 builds the shipped SimpleDiscriminator configuration only")
        self.model = nn.ModuleList([nn.Conv2d(1, 16, 4, 2), nn.Identity(), nn.Conv2d(16, 32, 4, 2), nn.Identity(),
                                    nn.Conv2d(32, 1, 1)])
        self.tail = nn.ModuleList([nn.Identity(), nn.Linear(62 * 62, 1, bias=False)])

    def forward(self, x):
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise NotImplementedError("uncltmo_b200 discriminator backward is not built yet: call under torch.no_grad()")
        if x.dim() != 4 or tuple(x.shape[1:]) != (1, 256, 256):
            raise ValueError("SimpleDiscriminator expects [N,1,256,256]")
        x = x.contiguous().float()
        n = x.shape[0]
        h = torch.empty((n, 16, 127, 127), device=x.device, dtype=torch.float32)
        fea = torch.empty((n, 1, 62, 62), device=x.device, dtype=torch.float32)
        logits = torch.empty((n, 1), device=x.device, dtype=torch.float32)
        m = self.model
        call("uncl_disc_forward", x, m[0].weight.detach().contiguous(), m[0].bias.detach(), m[2].weight.detach().contiguous(),
             m[2].bias.detach(), m[4].weight.detach().contiguous(), m[4].bias.detach(),
             self.tail[1].weight.detach().contiguous(), h, fea, logits, n, 256, 256)
        mean, con = plane_mean_contrast(fea)
        return logits, torch.cat([mean, con], dim=1)[:, :, None, None]
