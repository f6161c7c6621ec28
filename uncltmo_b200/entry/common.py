"""Shared helpers of the two entry points: settings, checkpoint loading, HDR file decode, PNG write.

Radiance .hdr frames are decoded ON THE DEVICE (read_hdr_image_device: the file's bytes are uploaded, the host only walks the
run-length packet headers); .exr goes through cv2 on the host, .npy is loaded directly.  The PNG write stays
on the host (cv2): its entropy-coding stage is sequential, see DESIGN.md section 8.  Everything between the decoded float
frame and the 8-bit result runs on the GPU.
"""
import os

import numpy as np
import torch

from ..frame import FramePipeline

EXTENSIONS = [".hdr", ".dng", ".exr", ".npy"]

# shipped hyper-parameters (both run_settings.npy files of the reference hold these, SURVEY.md §5 "config")
DEFAULT_MODEL_PARAMS = dict(model="unet", filters=32, depth=4, con_operator="square_and_square_root", last_layer="sigmoid",
                            unet_norm="none", stretch_g="none", g_doubleConvTranspose=1, bilinear=0, padding="replicate",
                            convtranspose_kernel=2, up_mode=0, add_frame=0, input_dim=1, factor_coeff=0.1,
                            final_shape_addition=0)


def get_model_params(model_name, train_settings_path="none"):
    """utils/model_save_util.py:620-652: hyper-parameters come from the pickled `vars(opt)` of the training run."""
    params = dict(DEFAULT_MODEL_PARAMS)
    if train_settings_path and os.path.exists(train_settings_path):
        saved = np.load(train_settings_path, allow_pickle=True)[()]
        rename = {"unet_depth": "depth", "use_contrast_ratio_f": None}
        for k, v in saved.items():
            k2 = rename.get(k, k)
            if k2 in params:
                params[k2] = v
    params["model_name"] = model_name
    return params


def get_layer_factor(con_operator):
    """test_imageTMO.py:105-114."""
    table = {"original_unet": 2, "square": 3, "square_root": 3, "gamma": 3, "square_and_square_root": 4,
             "square_and_square_root_manual_d": 4}
    if con_operator not in table:
        assert 0, "Unsupported con_operator request: {}".format(con_operator)
    return table[con_operator]


def set_parallel_net(net, device_):
    """test_imageTMO.py:117-122 wraps the net in nn.DataParallel; here frames / tile chains are sharded over one
    process per GPU instead (uncltmo_b200.dist), so the single-process net is returned unchanged."""
    return net


def create_G_net(generator_cls, model_params, device_, is_checkpoint, activation, output_dim, precision="bf16"):
    """test_imageTMO.py:86-102 / test_videoTMO.py:95-111."""
    layer_factor = get_layer_factor(model_params["con_operator"])
    if model_params["model"] != "unet":
        assert 0, "Unsupported g model request: {}".format(model_params["model"])
    return generator_cls(model_params["input_dim"], output_dim, model_params["last_layer"], depth=model_params["depth"],
                         layer_factor=layer_factor, con_operator=model_params["con_operator"],
                         filters=model_params["filters"], bilinear=model_params["bilinear"], network=model_params["model"],
                         dilation=0, to_crop=model_params["add_frame"], unet_norm=model_params["unet_norm"],
                         stretch_g=model_params["stretch_g"], activation=activation,
                         doubleConvTranspose=model_params["g_doubleConvTranspose"], padding_mode=model_params["padding"],
                         convtranspose_kernel=model_params["convtranspose_kernel"], up_mode=model_params["up_mode"],
                         precision=precision).to(device_)


def load_g_model(generator_cls, model_params, device, net_path, precision="bf16"):
    """test_imageTMO.py:74-83: checkpoint['modelG_state_dict'], 'module.' prefix stripped (model_save_util.py:188-198)."""
    g = create_G_net(generator_cls, model_params, device, is_checkpoint=True, activation="relu", output_dim=1,
                     precision=precision)
    checkpoint = torch.load(net_path, map_location=device)
    sd = {k[7:] if k.startswith("module.") else k: v for k, v in checkpoint["modelG_state_dict"].items()}
    g.load_state_dict(sd)
    g = set_parallel_net(g, device)
    g.eval()
    return g


def read_hdr_image(path):
    """utils/hdr_image_util.py:35-53: float32 RGB, HWC."""
    if path.endswith(".npy"):
        im = np.load(path, allow_pickle=True)
        if im.dtype == object:
            im = im[()]
            im = im.get("hdr_image", im) if isinstance(im, dict) else im
        return np.asarray(im, dtype=np.float32)
    if path.lower().endswith(".dng"):
        # the reference reads camera RAW through imageio's FreeImage plugin (hdr_image_util.py:35-53); neither imageio nor
        # a RAW decoder is a dependency here and OpenCV cannot decode DNG: fail with the reason instead of "cannot decode"
        raise IOError("%s: .dng (camera RAW) needs imageio + FreeImage as in the reference; convert to .hdr / .exr / .npy" % path)
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")     # recent OpenCV builds refuse .exr without it
    import cv2
    im = cv2.imread(path, cv2.IMREAD_ANYDEPTH | cv2.IMREAD_COLOR)
    if im is None:
        raise IOError("cannot decode %s (OpenCV: Radiance .hdr and OpenEXR are supported)" % path)
    return np.ascontiguousarray(im[..., ::-1], dtype=np.float32)


def read_hdr_image_device(path, device):
    """Radiance .hdr -> fp32 RGB planes [3,H,W] ON THE DEVICE: the file's bytes are uploaded as they are, the host only walks
    the run-length packet headers to find the scanline starts (uncl_hdr_scan_host), expansion and RGBE -> float run on
    the GPU (uncl_hdr_decode).  Bit-identical to cv2.imread(..., IMREAD_ANYDEPTH) (tests/test_gpu_entry.py)."""
    import ctypes
    import re
    from .. import _lib
    raw = np.fromfile(path, dtype=np.uint8)
    head = bytes(raw[:4096])
    m = re.search(rb"\n\n(-Y|\+Y) (\d+) (\+X|-X) (\d+)\n", head)
    if not head.startswith(b"#?") or m is None or m.group(1) != b"-Y" or m.group(3) != b"+X":
        raise IOError("%s: not a top-down Radiance RGBE file" % path)
    h, w = int(m.group(2)), int(m.group(4))
    offsets = np.empty(h + 1, dtype=np.int64)
    rle = ctypes.c_int(0)
    lib = _lib.lib()
    rc = lib.uncl_hdr_scan_host(raw.ctypes.data, raw.size, m.end(), w, h, offsets.ctypes.data, ctypes.addressof(rle))
    if rc != 0:
        raise IOError("%s: %s" % (path, lib.uncl_last_error().decode()))
    data = torch.from_numpy(raw).pin_memory().to(device, non_blocking=True)
    offs = torch.from_numpy(offsets).pin_memory().to(device, non_blocking=True)
    out = torch.empty((3, h, w), device=device, dtype=torch.float32)
    _lib.call("uncl_hdr_decode", data, offs, w, h, int(rle.value), out)
    return out


def load_lambda(f_factor_path, key):
    data = np.load(f_factor_path, allow_pickle=True)[()]
    return float(data[key])


def save_png(u8_hwc, output_path, name):
    import cv2
    os.makedirs(output_path, exist_ok=True)
    path = os.path.join(output_path, name + ".png")
    cv2.imwrite(path, np.ascontiguousarray(u8_hwc[..., ::-1]))
    print("result was saved to [%s]" % path)
    return path


def make_pipeline(g, model_params, overlap=64):
    return FramePipeline(g, factor_coeff=model_params["factor_coeff"], overlap=overlap)
