"""Image tone-mapping entry point; mirrors activate_trained_model/test_imageTMO.py (same flags and functions).

    python -m uncltmo_b200.entry.test_imageTMO --input_images_path input_images --model_path <dir with
        net_epoch5_iter62.pth + run_settings.npy> --f_factor_path lambda_data/...npy --output_path output
"""
import argparse
import os
import time

import numpy as np
import torch

from ..generator import UNet as _Generator
from . import common
from .common import get_layer_factor, set_parallel_net  # noqa: F401  (part of the mirrored surface)

extensions = common.EXTENSIONS

default_params = {"model_path": "model_weights_retrain20220815",
                  "model_name": "11_08_lr15D_size268_D_[1,1,1]_pad_0_G_ssr_doubleConvT__d1.0_struct_1.0[1,1,1]__trans2_replicate__noframe__min_log_0.1hist_fit_",
                  "input_images_path": "input_images",
                  "f_factor_path": "lambda_data/input_images_lambdas_HDRSdataset.npy",
                  "output_path": "output",
                  "mean_hist_path": "lambda_data/ldr_avg_hist_900_images_20_bins.npy",
                  "lambda_output_path": "lambda_data",
                  "bins": 20}


def get_args(argv=None):
    parser = argparse.ArgumentParser(description="Parser for gan network")
    for k in ("model_name", "input_images_path", "output_path", "model_path", "f_factor_path", "mean_hist_path",
              "lambda_output_path", "bins"):
        parser.add_argument("--" + k, type=str, default=default_params[k])
    # additions of this build (the reference hard-codes them: model_save_util.py:224-226, 303-304)
    parser.add_argument("--scale", type=int, default=4, help="host down-scale before tone mapping (reference: 4)")
    parser.add_argument("--overlap", type=int, default=64, help="tile overlap: 64 (1/4 res) or 192 (full res)")
    parser.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    return parser.parse_args(argv)


def run_trained_model(args):
    start = time.time()
    net_path = os.path.join(args.model_path, "net_epoch5_iter62.pth")
    train_settings_path = os.path.join(args.model_path, "run_settings.npy")
    model_params = common.get_model_params(args.model_name, train_settings_path)
    os.makedirs(args.output_path, exist_ok=True)
    if not torch.cuda.is_available():
        raise RuntimeError("uncltmo_b200 needs a CUDA device: there is no CPU path (run the reference for that)")
    device = torch.device("cuda")
    run_model_on_path(model_params, device, net_path, args.input_images_path, args.output_path, args.f_factor_path, None,
                      model_params["final_shape_addition"], scale=args.scale, overlap=args.overlap,
                      precision=args.precision)
    print("tone mapping took [%.2f] seconds" % (time.time() - start))


def run_model_on_path(model_params, device, net_path, input_images_path, output_images_path, f_factor_path, net_G,
                      final_shape_addition, scale=4, overlap=64, precision="bf16"):
    """test_imageTMO.py:58-71."""
    if not net_G:
        net_G = load_g_model(model_params, device, net_path, precision)
    print("\nModel [%s] was loaded successfully\n" % model_params["model"])
    pipe = common.make_pipeline(net_G, model_params, overlap)
    for img_name in sorted(os.listdir(input_images_path)):
        im_path = os.path.join(input_images_path, img_name)
        print("processing [%s]" % img_name)
        stem, ext = os.path.splitext(img_name)
        if ext not in extensions:
            raise Exception("invalid hdr file format: %s" % ext)
        run_model_on_single_image2(pipe, im_path, device, stem, output_images_path, f_factor_path, scale)


def run_model_on_single_image2(pipe, im_path, device, im_name, output_path, f_factor_path, scale=4):
    """utils/model_save_util.py:293-407: host decode (+ the hard-coded cv2 down-scale), then the GPU frame path."""
    lam = common.load_lambda(f_factor_path, im_name)
    if scale == 1 and im_path.lower().endswith(".hdr"):
        x = common.read_hdr_image_device(im_path, device)      # file bytes -> GPU, scanlines expanded there
    else:
        rgb = common.read_hdr_image(im_path)
        if scale != 1:
            import cv2
            rgb = cv2.resize(rgb, (rgb.shape[1] // scale, rgb.shape[0] // scale))
        x = torch.from_numpy(np.ascontiguousarray(rgb.transpose(2, 0, 1))).pin_memory().to(device, non_blocking=True)
    with torch.no_grad():
        u8 = pipe.tonemap(x, lam, uint8=True)
    return common.save_png(u8.cpu().numpy(), output_path, im_name + "_UnCLTMO")


def load_g_model(model_params, device, net_path, precision="bf16"):
    return common.load_g_model(_Generator, model_params, device, net_path, precision)


def create_G_net(model_params, device_, is_checkpoint, activation, output_dim, precision="bf16"):
    return common.create_G_net(_Generator, model_params, device_, is_checkpoint, activation, output_dim, precision)


if __name__ == "__main__":
    run_trained_model(get_args())
