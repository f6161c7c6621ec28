"""Oracle: the dataset loaders' per-crop processing.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference: utils/ProcessedDatasetFolderImg.py:44-168 (npy_loader), :13-22 (get_ldr_im); video variant
utils/ProcessedDatasetFolder.py:96-160.  cv2 does the resize / colour conversion exactly as in the reference.
"""
import numpy as np
import torch


def get_ldr_im(normalization, im, max_stretch, min_stretch):
    # ProcessedDatasetFolderImg.py:13-22 (torch tensor in, as in the reference)
    if normalization == "max_normalization":
        return im / im.max()
    if normalization == "bugy_max_normalization":
        return im / 255
    if normalization == "stretch":
        im = ((im - im.min()) / im.max()) * max_stretch - min_stretch
        return np.clip(im, 0, 1)
    return im


def crop_sample(color_im, rh, rw, xx, yy, hdr_mode, brightness_factor=None, normalization="max_normalization",
                max_stretch=1.0, min_stretch=0.0, use_ipp=True):
    """One crop of `npy_loader` for given draws.  color_im [H,W,3] float32.
    Returns (input_im [1,256,256], color [3,256,256], gray_norm, gray_shift) torch tensors (last two None in LDR mode).

    use_ipp=False runs cv2.resize through OpenCV's own (generic) bilinear code instead of the Intel IPP routine this
    image's cv2 build dispatches to; the two differ by ~4e-6 rel-L2 on float32 images.  The reference gets whichever
    its cv2 build has; the CUDA kernel implements OpenCV's algorithm."""
    import cv2
    color_im = np.asarray(color_im, np.float32)
    if (rh, rw) != color_im.shape[:2]:
        had = cv2.ipp.useIPP()
        cv2.ipp.setUseIPP(bool(use_ipp) and had)
        try:
            color_im = cv2.resize(color_im, (rw, rh))
        finally:
            cv2.ipp.setUseIPP(had)
    if color_im.shape[0] != 256 or color_im.shape[1] != 256:
        color_im = color_im[yy:yy + 256, xx:xx + 256, :]
    yuv = cv2.cvtColor(np.ascontiguousarray(color_im), cv2.COLOR_RGB2YUV)
    input_im = torch.from_numpy(np.ascontiguousarray(yuv[:, :, :1].transpose(2, 0, 1)))
    color = torch.from_numpy(np.ascontiguousarray(color_im.transpose(2, 0, 1)))
    if not hdr_mode:
        return get_ldr_im(normalization, input_im, max_stretch, min_stretch), color, None, None
    gray = (0.299 * color[0] + 0.587 * color[1] + 0.114 * color[2])[None]   # hdr_image_util.to_gray_tensor
    gray_norm = gray / gray.max()
    gray = gray - gray.min()
    a = torch.log10((gray / gray.max()) * brightness_factor + 1)
    return a / a.max(), color, gray_norm, gray
