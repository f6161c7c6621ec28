"""Lambda search objective on the GPU (utils/adaptive_lambda.py:7-21), batched over a population of candidates.

`cross_entropy(factors, gray_im, targets, bins)` takes a vector of candidate lambdas and returns one objective value
per candidate, so a population-based optimiser (the reference uses scipy differential_evolution, :62-63) evaluates a
whole generation in one launch instead of one numpy pass over the image per candidate.
"""
import numpy as np
import torch

from ._lib import call


def cross_entropy(factors, gray_im, targets, bins_):
    """factors: sequence of L lambdas; gray_im: CUDA fp32 image (any shape); targets: [bins] -> np.ndarray [L]."""
    if not (torch.is_tensor(gray_im) and gray_im.is_cuda):
        raise RuntimeError("uncltmo_b200.adaptive_lambda needs the image on a CUDA device (no CPU path)")
    g = gray_im.contiguous().float().reshape(-1)
    lam = torch.as_tensor(np.atleast_1d(np.asarray(factors, dtype=np.float64)), device=g.device)
    t = torch.as_tensor(np.asarray(targets, dtype=np.float32), device=g.device)
    bins_ = int(bins_)
    out = torch.empty(lam.numel(), device=g.device, dtype=torch.float32)
    ws = torch.empty((lam.numel() * bins_ + 4) * 4, device=g.device, dtype=torch.uint8)
    call("uncl_lambda_cross_entropy", g, g.numel(), lam, lam.numel(), t, bins_, out, ws)
    return out.cpu().numpy()


def calc_lambda_for_image(gray_im, targets, bins, bounds=(1.0, 1e9), maxiter=1000, seed=0):
    """differential evolution over lambda with the vectorised objective (adaptive_lambda.py:62-63)."""
    import scipy.optimize as optimize
    g = gray_im / gray_im.max()
    sol = optimize.differential_evolution(lambda pop: cross_entropy(pop.reshape(-1), g, targets, bins), bounds=[bounds],
                                          maxiter=maxiter, vectorized=True, updating="deferred", seed=seed)
    return float(sol.x[0]), float(sol.fun)
