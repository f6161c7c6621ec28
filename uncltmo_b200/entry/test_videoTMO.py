"""Video tone-mapping entry point; mirrors activate_trained_model/test_videoTMO.py (same flags and functions).

One sub-directory of --input_images_path per scene; lambda is looked up by scene name
(utils/model_save_util.py:248-249); all frames of a scene go through the recurrent generator as one clip.
"""
import argparse
import os
import time

import numpy as np
import torch

from ..generator import UNetVideo as _Generator
from . import common
from .common import get_layer_factor, set_parallel_net  # noqa: F401

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
    parser.add_argument("--overlap", type=int, default=64)
    parser.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    parser.add_argument("--max_frames", type=int, default=0, help="0 = whole scene (reference behaviour)")
    return parser.parse_args(argv)


def run_trained_model(args):
    start = time.time()
    net_path = os.path.join(args.model_path, "net_epoch10_iter124.pth")
    model_params = common.get_model_params(args.model_name, os.path.join(args.model_path, "run_settings.npy"))
    os.makedirs(args.output_path, exist_ok=True)
    if not torch.cuda.is_available():
        raise RuntimeError("uncltmo_b200 needs a CUDA device: there is no CPU path (run the reference for that)")
    run_model_on_path(model_params, torch.device("cuda"), net_path, args.input_images_path, args.output_path,
                      args.f_factor_path, None, model_params["final_shape_addition"], overlap=args.overlap,
                      precision=args.precision, max_frames=args.max_frames)
    print("tone mapping took [%.2f] seconds" % (time.time() - start))


def run_model_on_path(model_params, device, net_path, input_images_path, output_images_path, f_factor_path, net_G,
                      final_shape_addition, overlap=64, precision="bf16", max_frames=0):
    """test_videoTMO.py:58-80: one clip per scene directory."""
    if not net_G:
        net_G = load_g_model(model_params, device, net_path, precision)
    print("\nModel [%s] was loaded successfully\n" % model_params["model"])
    pipe = common.make_pipeline(net_G, model_params, overlap)
    for scene in sorted(os.listdir(input_images_path)):
        scene_dir = os.path.join(input_images_path, scene)
        names = sorted(n for n in os.listdir(scene_dir) if os.path.splitext(n)[1] in extensions)
        if max_frames:
            names = names[:max_frames]
        print("processing scene [%s]: %d frames" % (scene, len(names)))
        run_model_on_video(pipe, [os.path.join(scene_dir, n) for n in names], device,
                           [os.path.splitext(n)[0] for n in names], os.path.join(output_images_path, scene),
                           f_factor_path, scene)


def run_model_on_video(pipe, im_paths, device, im_names, output_path, f_factor_path, scene):
    """utils/model_save_util.py:567-614."""
    lam = common.load_lambda(f_factor_path, scene)
    if all(p.lower().endswith(".hdr") for p in im_paths):
        x = torch.stack([common.read_hdr_image_device(p, device) for p in im_paths])     # decoded on the GPU
    else:
        frames = np.stack([common.read_hdr_image(p).transpose(2, 0, 1) for p in im_paths])
        x = torch.from_numpy(np.ascontiguousarray(frames)).pin_memory().to(device, non_blocking=True)
    with torch.no_grad():
        u8 = pipe.tonemap_clip(x, lam, uint8=True).cpu().numpy()
    return [common.save_png(u8[i], output_path, im_names[i] + "_UnCLTMO") for i in range(len(im_paths))]


def load_g_model(model_params, device, net_path, precision="bf16"):
    return common.load_g_model(_Generator, model_params, device, net_path, precision)


def create_G_net(model_params, device_, is_checkpoint, activation, output_dim, precision="bf16"):
    return common.create_G_net(_Generator, model_params, device_, is_checkpoint, activation, output_dim, precision)


if __name__ == "__main__":
    run_trained_model(get_args())
