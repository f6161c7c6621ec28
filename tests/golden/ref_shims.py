"""Shims that make the upstream reference importable on CPU in THIS container.

Only `tests/golden/make_golden.py` uses this (it cannot run on the GPU box: /root/reference
does not exist there).  It installs throw-away stand-ins for the third-party modules the
reference imports but the image lacks and exposes the reference
modules under their own names.  No reference code is copied: it is imported in place.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("UNCLTMO_REFERENCE", "/root/reference")


def _mod(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _DropPath(torch.nn.Module):
    """timm.models.layers.DropPath semantics: per-sample Bernoulli keep, 1/keep scale, identity in eval."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x * mask / keep


def install():
    if not os.path.isdir(REF):
        raise RuntimeError("reference tree not present at %s" % REF)
    if "timm" in sys.modules and getattr(sys.modules["timm"], "_uncl_stub", False):
        return
    if not hasattr(np, "float"):
        np.float = float  # gcn_lib/pos_embed.py:74 uses the removed alias
    t = _mod("timm", _uncl_stub=True)
    _mod("timm.data", IMAGENET_DEFAULT_MEAN=(0.485, 0.456, 0.406), IMAGENET_DEFAULT_STD=(0.229, 0.224, 0.225))
    _mod("timm.models")
    _mod("timm.models.helpers", load_pretrained=lambda *a, **k: None)
    _mod("timm.models.registry", register_model=lambda f: f)
    _mod("timm.models.layers", DropPath=_DropPath, to_2tuple=lambda x: (x, x),
         trunc_normal_=torch.nn.init.trunc_normal_)
    import cv2

    def _imread(path, *a, **k):
        im = cv2.imread(path, cv2.IMREAD_ANYDEPTH | cv2.IMREAD_COLOR)
        return im[..., ::-1].copy()

    _mod("imageio", imread=_imread, imwrite=lambda *a, **k: None,
         plugins=types.SimpleNamespace(freeimage=types.SimpleNamespace(download=lambda: None)))
    _mod("matplotlib", use=lambda *a, **k: None)
    _mod("matplotlib.pyplot")
    sk = _mod("skimage")
    _mod("skimage.transform", resize=None)
    _mod("skimage.io")
    _mod("skimage.exposure")
    _mod("skimage.color")

    def _view_as_blocks(arr, block_shape):
        bh, bw = block_shape
        h, w = arr.shape
        return arr.reshape(h // bh, bh, w // bw, bw).transpose(0, 2, 1, 3)

    _mod("skimage.util", view_as_blocks=_view_as_blocks)
    sk.transform = sys.modules["skimage.transform"]
    sk.util = sys.modules["skimage.util"]
    _mod("torchsummary", summary=lambda *a, **k: None)
    _mod("contracts", contract=lambda *a, **k: (lambda f: f))
    _mod("wget")
    import scipy.signal
    import scipy.signal.windows
    if not hasattr(scipy.signal, "gaussian"):
        scipy.signal.gaussian = scipy.signal.windows.gaussian
    if REF not in sys.path:
        sys.path.insert(0, REF)


def reference_modules():
    """Return the reference modules on the hot path, imported in place."""
    install()
    import importlib
    names = dict(
        gen_img="models.unet_multi_filters.Unet_singleFrame",
        gen_vid="models.unet_multi_filters.Unet",
        disc="models.Discriminator",
        struct_loss="models.struct_loss",
        hdr_util="utils.hdr_image_util",
        dl_util="utils.data_loader_util",
        save_util="utils.model_save_util",
        params="utils.params",
    )
    out = {}
    for k, v in names.items():
        out[k] = importlib.import_module(v)
    return types.SimpleNamespace(**out)
