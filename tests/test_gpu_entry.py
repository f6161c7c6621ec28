"""Entry points end to end on the GPU (activate_trained_model/test_imageTMO.py:58-71,
test_videoTMO.py:58-80): a directory of .npy / .hdr frames through run_model_on_path -> PNG files, against the oracle's frame
path; and the wrappers the reference puts around its modules (nn.DataParallel in set_parallel_net, DDP-style wrapping)."""
import os

import numpy as np
import pytest
import torch

import oracle
from uncltmo_b200 import synth
from uncltmo_b200.entry import common, test_imageTMO as entry_img, test_videoTMO as entry_vid
from uncltmo_b200.generator import UNet, UNetVideo
from uncltmo_b200.weights import make_generator_state_dict

pytestmark = pytest.mark.gpu
G_ARGS = (1, 1, "sigmoid", 4, 4, "square_and_square_root", 32, 0, "unet", 0, 0, "none", "none", "relu", 1, "replicate", 2)


def test_image_entry_on_a_directory(tmp_path):
    import cv2
    inp, out = tmp_path / "in", tmp_path / "out"
    inp.mkdir()
    frames = {"a": synth.hdr_frame(268, 300, seed=31), "b": synth.hdr_frame(300, 268, seed=32)}
    np.save(inp / "a.npy", frames["a"].transpose(1, 2, 0))                                  # HWC float32, as the loaders store
    cv2.imwrite(str(inp / "b.hdr"), np.ascontiguousarray(frames["b"].transpose(1, 2, 0)[..., ::-1]))   # Radiance RGBE
    lam = {"a": 50.0, "b": 120.0}
    np.save(tmp_path / "lambdas.npy", lam, allow_pickle=True)
    sd = make_generator_state_dict()
    net = UNet(*G_ARGS, up_mode=0, precision="fp32").cuda().eval()
    net.load_state_dict(sd)
    params = common.get_model_params("unit-test")
    entry_img.run_model_on_path(params, torch.device("cuda"), None, str(inp), str(out), str(tmp_path / "lambdas.npy"), net,
                                params["final_shape_addition"], scale=1, overlap=64, precision="fp32")
    for name in ("a", "b"):
        png = cv2.imread(str(out / (name + "_UnCLTMO.png")))[..., ::-1]
        rgb = torch.from_numpy(common.read_hdr_image(str(inp / (name + (".npy" if name == "a" else ".hdr")))).transpose(2, 0, 1).copy())
        want = oracle.frame_path.to_uint8_stretch(oracle.tonemap_frame(rgb, lam[name], lambda t: oracle.unet_forward(sd, t)[0]))
        assert png.shape == want.shape
        assert np.abs(png.astype(int) - want.astype(int)).max() <= 1, name
    with pytest.raises(Exception):
        (inp / "c.txt").write_text("x")
        entry_img.run_model_on_path(params, torch.device("cuda"), None, str(inp), str(out), str(tmp_path / "lambdas.npy"), net, 0, scale=1)


def test_video_entry_on_a_scene(tmp_path):
    inp, out = tmp_path / "in", tmp_path / "out"
    (inp / "scene1").mkdir(parents=True)
    clip = synth.hdr_clip(3, 268, 300, seed=33)
    for k in range(3):
        np.save(inp / "scene1" / ("f%03d.npy" % k), clip[k].transpose(1, 2, 0))
    np.save(tmp_path / "lambdas.npy", {"scene1": 371.4}, allow_pickle=True)
    net = UNetVideo(*G_ARGS, up_mode=0, precision="bf16").cuda().eval()
    net.load_state_dict(make_generator_state_dict())
    params = common.get_model_params("unit-test")
    entry_vid.run_model_on_path(params, torch.device("cuda"), None, str(inp), str(out), str(tmp_path / "lambdas.npy"), net, 0)
    files = sorted(os.listdir(out / "scene1"))
    assert files == ["f000_UnCLTMO.png", "f001_UnCLTMO.png", "f002_UnCLTMO.png"]


def test_checkpoint_round_trip_with_module_prefix(tmp_path):
    """load_g_model: checkpoint['modelG_state_dict'] saved from a DataParallel-wrapped net ('module.' prefix),
    utils/model_save_util.py:188-198."""
    sd = make_generator_state_dict()
    torch.save({"modelG_state_dict": {"module." + k: v for k, v in sd.items()}}, tmp_path / "net.pth")
    params = common.get_model_params("unit-test")
    g = entry_img.load_g_model(params, torch.device("cuda"), str(tmp_path / "net.pth"), precision="bf16")
    assert not g.training
    for k, v in g.state_dict().items():
        assert torch.equal(v.cpu(), sd[k]), k


def test_data_parallel_and_ddp_style_wrapping():
    """nn.DataParallel (the reference's set_parallel_net, test_imageTMO.py:117-122) on one device, and the
    `module.`-prefixed state_dict it produces; a single-process DDP wrap when a process group can be created."""
    sd = make_generator_state_dict()
    net = UNet(*G_ARGS, up_mode=0, precision="bf16").cuda().eval()
    net.load_state_dict(sd)
    x = torch.from_numpy(synth.normalised_batch(2, seed=2)).cuda()
    with torch.no_grad():
        want, _ = net(x)
        dp = torch.nn.DataParallel(net, device_ids=[0])
        got, feats = dp(x)
    assert torch.equal(got, want) and feats.shape == (2, 32, 256, 256)
    assert all(k.startswith("module.") for k in dp.state_dict())
    import torch.distributed as dist
    if not dist.is_initialized():
        import socket
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=0, world_size=1)
        created = True
    else:
        created = False
    try:
        net.train()
        net.drop_path_prob = 0.0
        ddp = torch.nn.parallel.DistributedDataParallel(net, device_ids=[0])
        out, _ = ddp(x)
        out.sum().backward()
        assert all(p.grad is not None for p in net.parameters() if p.requires_grad)
    finally:
        if created:
            dist.destroy_process_group()


@pytest.mark.parametrize("h,w", [(268, 300), (1080, 1920), (37, 7)])
def test_hdr_decode_on_device_matches_opencv(tmp_path, h, w):
    """Radiance .hdr decode with the scanlines expanded on the GPU (uncl_hdr_scan_host + uncl_hdr_decode) is bit-identical
    to cv2.imread: run-length coded files as OpenCV writes them, and flat RGBE (widths < 8 cannot be run-length coded)."""
    import cv2
    rgb = synth.hdr_frame(h, w, seed=41)
    rgb[:, : h // 3] *= 1e-3            # long runs of equal exponents and a wide dynamic range
    rgb[:, -2:] = 0.0                   # zero pixels (exponent byte 0)
    path = str(tmp_path / "f.hdr")
    assert cv2.imwrite(path, np.ascontiguousarray(rgb.transpose(1, 2, 0)[..., ::-1]))
    want = cv2.imread(path, cv2.IMREAD_ANYDEPTH | cv2.IMREAD_COLOR)[..., ::-1].transpose(2, 0, 1)
    got = common.read_hdr_image_device(path, torch.device("cuda"))
    torch.cuda.synchronize()
    assert got.shape == (3, h, w)
    assert np.array_equal(got.cpu().numpy(), want)
    # a corrupted stream is refused by the host scan, not decoded into garbage
    raw = bytearray(open(path, "rb").read())
    if w >= 8:
        raw[-1:] = b""
        bad = str(tmp_path / "bad.hdr")
        open(bad, "wb").write(bytes(raw))
        with pytest.raises(IOError):
            common.read_hdr_image_device(bad, torch.device("cuda"))
