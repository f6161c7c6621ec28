"""models/Blocks.py output activations / normalisers (off-path in the shipped configuration), forward."""
import torch
import torch.nn as nn

from ._lib import call

_MODES = {"Exp": 0, "MySig": 1, "Clip": 2, "MaxNormalization": 3, "MaxNormalizationEpsilon": 4,
          "BatchMaxNormalization": 5, "MinMaxNormalization": 6}


def _apply(x, mode, param=0.0):
    if not x.is_cuda:
        raise RuntimeError("uncltmo_b200 has no CPU path: move the input to a CUDA device")
    x = x.contiguous().float()
    n = x.shape[0]
    out = torch.empty_like(x)
    scratch = torch.empty(2 * n, device=x.device, dtype=torch.float32)
    call("uncl_blocks_apply", x, out, n, x.numel() // n, mode, float(param), scratch)
    return out


def _make(name, has_factor=False):
    class _Block(nn.Module):
        def __init__(self, factor=0.0):
            super().__init__()
            self.factor = factor

        def forward(self, x):
            return _apply(x, _MODES[name], self.factor)

    _Block.__name__ = _Block.__qualname__ = name
    return _Block


Exp = _make("Exp")
MySig = _make("MySig", True)
Clip = _make("Clip")
MaxNormalization = _make("MaxNormalization")
MaxNormalizationEpsilon = _make("MaxNormalizationEpsilon")
BatchMaxNormalization = _make("BatchMaxNormalization")
MinMaxNormalization = _make("MinMaxNormalization")
