"""Drop-in StructLoss (models/struct_loss.py:7-40): pyramid structural loss, forward on an sm_100a stencil kernel."""
import torch
import torch.nn as nn

from ._lib import call


class StructLoss(nn.Module):
    def __init__(self, pyramid_weight_list, window_size=5, pyramid_pow=False, use_c3=False, struct_method="gamma_struct",
                 crop_input=True, final_shape_addition=0):
        super().__init__()
        if window_size != 5 or final_shape_addition != 0:
            raise NotImplementedError("uncltmo_b200 builds the shipped StructLoss (5x5 window, no frame) only")
        self.pyramid_weight_list = pyramid_weight_list
        self.window_size = window_size

    def forward(self, fake, hdr_input_original_gray_norm, hdr_input, pyramid_weight_list):
        """The second argument is ignored, as in the reference (struct_loss.py:23,39)."""
        if torch.is_grad_enabled() and fake.requires_grad:
            raise NotImplementedError("uncltmo_b200 struct-loss backward is not built yet: call under torch.no_grad()")
        if fake.shape != hdr_input.shape or fake.dim() != 4 or fake.shape[1] != 1:
            raise ValueError("StructLoss expects two [N,1,H,W] tensors of equal shape")
        w = [float(v) for v in (pyramid_weight_list.tolist() if torch.is_tensor(pyramid_weight_list) else pyramid_weight_list)]
        n, _, h, wd = fake.shape
        fake = fake.contiguous().float()
        hdr = hdr_input.contiguous().float()
        import ctypes
        wh = (ctypes.c_float * len(w))(*w)
        scratch = torch.empty(2 * n * (h // 2) * (wd // 2) * 4 // 3 + 64, device=fake.device, dtype=torch.float32)
        out = torch.empty(1, device=fake.device, dtype=torch.float32)
        call("uncl_struct_loss_fwd", fake, hdr, n, h, wd, len(w), ctypes.cast(wh, ctypes.c_void_p).value, out, scratch)
        return out[0]
