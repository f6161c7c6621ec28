"""Oracle: the frame path around the generator.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

log-lambda normalisation, replicate-pad to the U-Net grid, 256x256 tiling with 64-px overlap and the
reference's sequential linear cross-fade, percentile clamp / stretch / back-to-colour / crop.

Reference: utils/model_save_util.py:219-240 (load_inference2), :242-263 (load_inference_testvideo),
:293-407 (run_model_on_single_image2), :409-486 (test_big_size_image2), :488-565 (5-D video variant),
:567-614 (run_model_on_video); utils/data_loader_util.py:135-185; utils/hdr_image_util.py:76-132, 93-102, 237-245.
"""
import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-8


def to_gray(rgb):
    # hdr_image_util.py:76-82
    return (0.299 * rgb[0] + 0.587 * rgb[1] + 0.114 * rgb[2])[None]


def log_lambda_normalise(rgb, lam, factor_coeff=0.1):
    """rgb [3,H,W] -> (rgb (shifted if negative), gray_log [1,H,W] in [0,1]).  model_save_util.py:232-239."""
    f = lam * 255 * factor_coeff
    if rgb.min() < 0:
        rgb = rgb - rgb.min()
    g = to_gray(rgb)
    g = g - g.min()
    g = torch.log10((g / g.max()) * f + 1)
    return rgb, g / g.max()


def resize_im(im):
    """Replicate-pad [C,H,W] to 16*floor(H/16)+16 (always pads).  data_loader_util.py:135-157, 175-179."""
    h, w = im.shape[1], im.shape[2]
    dy = 16 * (h // 16) + 16 - h
    dx = 16 * (w // 16) + 16 - w
    im = F.pad(im[None], (dx // 2, dx - dx // 2, dy // 2, dy - dy // 2), mode="replicate")[0]
    return im, dy, dx


def tile_grid(length, patch=256, overlap=64):
    """Tile start offsets along one axis: regular tiles while patch*i - overlap*(i-1) < length, then one
    tile anchored at length - patch.  model_save_util.py:416-440."""
    starts, i = [], 1
    while patch * i - overlap * (i - 1) < length:
        starts.append((patch - overlap) * (i - 1))
        i += 1
    if not starts:
        raise ValueError("tiling needs a padded extent > %d (got %d)" % (patch, length))
    return starts, length - patch


def _blend_row(get_patch, starts, last, width, patch=256, overlap=64):
    """One horizontal band: sequential left-to-right cross-fade; model_save_util.py:421-446."""
    band = None
    end = 0
    for j, s in enumerate(starts):
        p = get_patch(s)
        if band is None:
            band = torch.zeros(p.shape[:-1] + (width,), dtype=p.dtype)
            band[..., s:s + patch] = p
        else:
            for i in range(overlap):
                band[..., s + i] = band[..., s + i] * (overlap - 1 - i) / (overlap - 1) + p[..., i] * i / (overlap - 1)
            band[..., s + overlap:s + patch] = p[..., overlap:]
        end = s + patch
    p = get_patch(last)
    rng = end - last
    for i in range(rng):
        band[..., last + i] = band[..., last + i] * (rng - 1 - i) / (rng - 1) + p[..., i] * i / (rng - 1)
    band[..., end:] = p[..., rng:]
    return band


def tile_and_blend(x, model_fn, patch=256, overlap=64):
    """x [..., H, W] -> same shape.  model_fn maps a [..., 256, 256] tile to its output tile.

    Follows test_big_size_image2 (4-D) / test_big_size_image (5-D): row bands top to bottom, each band
    cross-faded left to right, bands cross-faded over the vertical overlap; last band/column anchored at the
    far edge and faded over however much it overlaps what is already there.
    """
    H, W = x.shape[-2], x.shape[-1]
    ys, ylast = tile_grid(H, patch, overlap)
    xs, xlast = tile_grid(W, patch, overlap)
    out = torch.zeros_like(x)
    end = 0
    for bi, y0 in enumerate(ys + [ylast]):
        band = _blend_row(lambda s: model_fn(x[..., y0:y0 + patch, s:s + patch]), xs, xlast, W, patch, overlap)
        if bi == 0:
            out[..., y0:y0 + patch, :] = band
            end = y0 + patch
        elif bi < len(ys):
            for i in range(overlap):
                out[..., y0 + i, :] = out[..., y0 + i, :] * (overlap - 1 - i) / (overlap - 1) + band[..., i, :] * i / (overlap - 1)
            out[..., y0 + overlap:y0 + patch, :] = band[..., overlap:, :]
            end = y0 + patch
        else:
            rng = end - ylast
            for i in range(rng):
                out[..., ylast + i, :] = out[..., ylast + i, :] * (rng - 1 - i) / (rng - 1) + band[..., i, :] * i / (rng - 1)
            out[..., end:, :] = band[..., rng:, :]
    return out


def back_to_color(rgb, fake):
    # hdr_image_util.py:122-132: (rgb / (Y + eps))^0.5 * fake
    if rgb.min() < 0:
        rgb = rgb - rgb.min()
    return torch.pow(rgb / (to_gray(rgb) + EPS), 0.5) * fake


def postprocess_frame(fake, rgb_pad, dy, dx):
    """fake [1,1,H1,W1], rgb_pad [3,H1,W1] -> colour LDR [3,H,W].  model_save_util.py:389-402."""
    arr = fake.numpy()
    hi = np.percentile(arr, 99.5)
    lo = np.percentile(arr, 0.5)
    f2 = fake.clamp(float(lo), float(hi))
    st = (f2 - f2.min()) / (f2.max() - f2.min())
    col = back_to_color(rgb_pad, st[0])
    im_max = col.max()
    col = col[:, dy // 2:-(dy - dy // 2), dx // 2:-(dx - dx // 2)]
    return col.clamp(min=0, max=float(im_max))


def to_uint8_stretch(col):
    """hdr_image_util.py:237-245 + :93-102: clamp(0,1), percentile(0.1, 99.0) stretch, clip, *255 -> uint8 HWC."""
    t = col.clamp(0, 1).permute(1, 2, 0).numpy()
    hi = np.percentile(t, 99.0)
    lo = np.percentile(t, 0.1)
    t = np.clip((t - lo) / (hi - lo), 0, 1)
    return (t * 255).astype("uint8")


def tonemap_frame(rgb, lam, model_fn, factor_coeff=0.1):
    """Whole image path of run_model_on_single_image2 minus file I/O and the hard-coded /4 resize."""
    rgb, g = log_lambda_normalise(rgb, lam, factor_coeff)
    rgb_p, dy, dx = resize_im(rgb)
    g_p, _, _ = resize_im(g)
    fake = tile_and_blend(g_p[None], model_fn)
    return postprocess_frame(fake, rgb_p, dy, dx)


def lambda_cross_entropy(factor, gray_im, targets, bins_):
    """utils/adaptive_lambda.py:7-21 (numpy)."""
    g = np.log10(np.asarray(gray_im) * factor + 1)
    g = g / g.max()
    pred, _ = np.histogram(g.reshape(-1), bins=bins_, density=True, range=(0, 1))
    return -np.sum(np.asarray(targets) * np.log(pred + 1e-9)) / pred.shape[0]
