"""Synthetic inputs of the shapes BASELINE.json names (there are no datasets in this environment).

Definitions follow SURVEY.md §8(d): log-normal HDR frames spanning ~5 decades, smoothed-uniform LDR
batches quantised to 8 bits, translated clips.  Everything is numpy + explicit seeds so the same arrays
are produced here, in the golden generator and on the GPU box.
"""
import numpy as np


def _box(a, k):
    """k x k box mean with edge replication, over the first two axes."""
    p = k // 2
    a = np.pad(a, ((p, p), (p, p)) + ((0, 0),) * (a.ndim - 2), mode="edge")
    c = np.cumsum(np.cumsum(a, axis=0, dtype=np.float64), axis=1)
    c = np.pad(c, ((1, 0), (1, 0)) + ((0, 0),) * (a.ndim - 2))
    return ((c[k:, k:] - c[:-k, k:] - c[k:, :-k] + c[:-k, :-k]) / (k * k)).astype(np.float32)


def hdr_frame(h, w, seed=0):
    """rgb [3,h,w] float32, strictly positive, ~5 decades of luminance."""
    rng = np.random.default_rng(seed)
    base = _box(rng.standard_normal((h, w)).astype(np.float32), 9)
    base = base / (base.std() + 1e-12)
    lum = np.exp(1.2 * base).astype(np.float32)
    return np.stack([lum * 1.0, lum * 0.9, lum * 0.8]).astype(np.float32)


def hdr_clip(t, h, w, seed=0):
    """[t,3,h,w]: frame k = frame 0 translated by (k, 2k) px (wrapping) + 1 % multiplicative noise."""
    rng = np.random.default_rng(seed + 1)
    f0 = hdr_frame(h, w, seed)
    out = []
    for k in range(t):
        fk = np.roll(f0, (k, 2 * k), axis=(1, 2))
        out.append(fk * (1.0 + 0.01 * rng.standard_normal((1, h, w)).astype(np.float32)))
    return np.stack(out).astype(np.float32)


def ldr_batch(n, h=256, w=256, seed=1):
    """[n,1,h,w] in [0,1], quantised to 1/255 like decoded 8-bit LDR crops."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        u = _box(rng.random((h, w)).astype(np.float32), 5)
        u = (u - u.min()) / (u.max() - u.min())
        out.append(np.round(u * 255.0) / 255.0)
    return np.stack(out)[:, None].astype(np.float32)


def normalised_batch(n, h=256, w=256, seed=2, lam=50.0, factor_coeff=0.1):
    """[n,1,h,w] log-lambda-normalised luminance crops (what the training loader hands the generator)."""
    out = []
    for i in range(n):
        rgb = hdr_frame(h, w, seed * 1000 + i)
        y = 0.299 * rgb[0] + 0.587 * rgb[1] + 0.114 * rgb[2]
        y = y - y.min()
        y = np.log10(y / y.max() * (lam * 255 * factor_coeff) + 1)
        out.append(y / y.max())
    return np.stack(out)[:, None].astype(np.float32)
