"""Drop-in StructLoss (models/struct_loss.py:7-40): pyramid structural loss on sm_100a stencil kernels (fwd + bwd)."""
import torch
import torch.nn as nn

from .autograd_losses import StructLossFn


class StructLoss(nn.Module):
    def __init__(self, pyramid_weight_list, window_size=5, pyramid_pow=False, use_c3=False, struct_method="gamma_struct",
                 crop_input=True, final_shape_addition=0):
        super().__init__()
        if window_size != 5 or final_shape_addition != 0:
            raise NotImplementedError("uncltmo_b200 builds the shipped StructLoss (5x5 window, no frame) only")
        self.pyramid_weight_list = pyramid_weight_list
        self.window_size = window_size
        self._w_cache = None

    def _weights(self, pyramid_weight_list):
        # the reference keeps the weights as a (device) tensor; read it once instead of once per step
        key = id(pyramid_weight_list)
        if self._w_cache is None or self._w_cache[0] != key:
            w = pyramid_weight_list.tolist() if torch.is_tensor(pyramid_weight_list) else list(pyramid_weight_list)
            self._w_cache = (key, [float(v) for v in w])
        return self._w_cache[1]

    def forward(self, fake, hdr_input_original_gray_norm, hdr_input, pyramid_weight_list):
        """The second argument is ignored, as in the reference (struct_loss.py:23,39)."""
        if fake.shape != hdr_input.shape or fake.dim() != 4 or fake.shape[1] != 1:
            raise ValueError("StructLoss expects two [N,1,H,W] tensors of equal shape")
        if not fake.is_cuda:
            raise RuntimeError("uncltmo_b200 has no CPU path: move the inputs to a CUDA device")
        return StructLossFn.apply(fake, hdr_input.detach(), self._weights(pyramid_weight_list))
