"""Seeded inputs shared by the golden generator (tests/golden/make_golden.py) and the tests.

Nothing here reads /root/reference; the arrays are rebuilt from seeds on whichever box runs the tests.
"""
import numpy as np
import torch

from uncltmo_b200 import synth

LAMBDA = 50.0


def generator_input():
    """[2,1,256,256] log-lambda-normalised luminance crops."""
    return torch.from_numpy(synth.normalised_batch(2, seed=2))


def video_input():
    """[1,2,1,256,256]: two consecutive frames of a translated clip, normalised per frame."""
    clip = synth.hdr_clip(2, 256, 256, seed=7)
    frames = []
    for k in range(2):
        y = 0.299 * clip[k, 0] + 0.587 * clip[k, 1] + 0.114 * clip[k, 2]
        y = y - y.min()
        y = np.log10(y / y.max() * (LAMBDA * 255 * 0.1) + 1)
        frames.append(y / y.max())
    return torch.from_numpy(np.stack(frames)[None, :, None].astype(np.float32))


def ldr_input():
    """[3,1,256,256] 8-bit-quantised LDR crops."""
    return torch.from_numpy(synth.ldr_batch(3, seed=1))


def logits_pair():
    rng = np.random.default_rng(11)
    return (torch.from_numpy(rng.standard_normal((16, 1)).astype(np.float32)),
            torch.from_numpy((rng.standard_normal((16, 1)) * 1.5 + 0.3).astype(np.float32)))


def nce_features_small():
    """Three [16,2,1,1] D-feature tensors."""
    rng = np.random.default_rng(12)
    return tuple(torch.from_numpy((rng.random((16, 2, 1, 1)) * s).astype(np.float32)) for s in (1.0, 0.8, 1.2))


def nce_features_map():
    """Three [3,32,64,64] feature maps (post-ReLU like)."""
    rng = np.random.default_rng(13)
    return tuple(torch.from_numpy(np.maximum(rng.standard_normal((3, 32, 64, 64)) * s, 0).astype(np.float32))
                 for s in (0.3, 0.25, 0.35))


def small_frame():
    """rgb [3,268,300]: pads to 272x304 -> 2x2 tiles."""
    return synth.hdr_frame(268, 300, seed=5)


def blend_field(h, w):
    rng = np.random.default_rng(h * 10007 + w)
    return rng.random((1, 1, h, w)).astype(np.float32)


def cheap_model(t):
    """Stand-in 'generator' for pinning the tiling arithmetic: depends on tile-local position so that
    overlapping tiles disagree (otherwise blending would be invisible)."""
    h, w = t.shape[-2], t.shape[-1]
    yy = torch.linspace(0, 1, h, dtype=t.dtype).view(*([1] * (t.ndim - 2)), h, 1)
    xx = torch.linspace(0, 1, w, dtype=t.dtype).view(*([1] * (t.ndim - 2)), 1, w)
    return 0.5 * t + 0.3 * yy * xx + 0.1 * xx


def act_stride(shape):
    """Spatial stride used when storing one image's activation [C,H,W] in the fixtures (<= ~40k values)."""
    import math
    return max(1, math.ceil(math.sqrt(shape[1] * shape[2] * shape[3] / 40000.0)))


def train_batch(b, video=False):
    """(hdr_input, real_ldr_pos, real_ldr_neg), each [b/2, 2, 1, 256, 256]: loader batch x 2 crops for the image trainer
    (ProcessedDatasetFolderImg.py), b/2 clips of T = 2 frames for the video trainer.  b = 16 is the reference recipe
    (run_imageTMO_train.sh:6-7: batch 8 x 2)."""
    hdr = torch.from_numpy(synth.normalised_batch(b, seed=4)).reshape(b // 2, 2, 1, 256, 256)
    pos = torch.from_numpy(synth.ldr_batch(b, seed=5)).reshape(b // 2, 2, 1, 256, 256)
    neg = torch.from_numpy(synth.ldr_batch(b, seed=6)).reshape(b // 2, 2, 1, 256, 256)
    return hdr, pos, neg


def droppath_masks(b, keep=0.95):
    """Two [b] per-sample scale vectors (mask / keep_prob) for the two residual branches of the graph block: sample 3 is
    dropped in the first branch, sample b-5 in the second (train-mode DropPath p = 0.05, Unet_singleFrame.py:62-77)."""
    m0, m1 = torch.full((b,), 1.0 / keep), torch.full((b,), 1.0 / keep)
    m0[3 % b] = 0.0
    m1[(b - 5) % b] = 0.0
    return [m0, m1]
