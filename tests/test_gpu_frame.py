"""GPU parity tests of the frame path (normalise, pad, tiling, blend, percentiles, post-process) via the C ABI."""
import numpy as np
import pytest
import torch

import golden_inputs as gi
import oracle
from uncltmo_b200 import synth
from uncltmo_b200.frame import FramePipeline
from uncltmo_b200.generator import UNet
from uncltmo_b200.weights import make_generator_state_dict

pytestmark = pytest.mark.gpu
G_ARGS = (1, 1, "sigmoid", 4, 4, "square_and_square_root", 32, 0, "unet", 0, 0, "none", "none", "relu", 1, "replicate", 2)


@pytest.fixture(autouse=True)
def _no_grad():
    """Inference tests run without autograd; tests that need it re-enable it locally."""
    with torch.no_grad():
        yield


def maxabs(a, b):
    return (torch.as_tensor(a).double().cpu() - torch.as_tensor(b).double().cpu()).abs().max().item()


class CheapG:
    """Stand-in generator so the blend is tested without the network in the way."""

    def tonemap_tiles(self, t):
        return gi.cheap_model(t.cpu()).cuda()


@pytest.fixture(scope="module")
def net():
    n = UNet(*G_ARGS, up_mode=0, precision="fp32").cuda().eval()
    n.load_state_dict(make_generator_state_dict())
    return n


def test_normalise_and_pad(golden):
    rgb = torch.from_numpy(gi.small_frame())
    _, gray = oracle.log_lambda_normalise(rgb, gi.LAMBDA)
    gp, dy, dx = oracle.resize_im(gray)
    pipe = FramePipeline(CheapG())
    got, _ = pipe.normalise_pad(rgb.cuda(), gi.LAMBDA)
    assert got.shape == gp.shape[1:]
    assert maxabs(got, gp[0]) <= 2e-6
    assert maxabs(got[dy // 2:dy // 2 + 268:2, dx // 2:dx // 2 + 300:2], golden["norm_gray_s2"][0]) <= 2e-6


def test_normalise_negative_input_shift():
    rgb = torch.from_numpy(gi.small_frame()) - 0.37  # exr-style negative values (model_save_util.py:233-234)
    _, gray = oracle.log_lambda_normalise(rgb, 371.4)
    gp, _, _ = oracle.resize_im(gray)
    got, _ = FramePipeline(CheapG()).normalise_pad(rgb.cuda(), 371.4)
    assert maxabs(got, gp[0]) <= 5e-6


@pytest.mark.parametrize("shape,overlap", [((272, 304), 64), ((464, 656), 64), ((784, 1040), 64), ((464, 464), 192)])
def test_gather_and_blend(shape, overlap):
    x = torch.from_numpy(gi.blend_field(*shape))
    ref = oracle.tile_and_blend(x, gi.cheap_model, overlap=overlap)
    pipe = FramePipeline(CheapG(), overlap=overlap)
    pl = pipe.plan(shape[0] - 16, shape[1] - 16, torch.device("cuda"))
    assert (pl.h1, pl.w1) == shape
    tiles = pipe.gather_tiles(x[0, 0].cuda(), pl)
    for t, (oy, ox) in enumerate(pl.origins.cpu().tolist()):
        assert torch.equal(tiles[t, 0].cpu(), x[0, 0, oy:oy + 256, ox:ox + 256])
    got = pipe.blend(pipe.run_generator(tiles), pl)
    assert maxabs(got, ref[0, 0]) <= 2e-6


def test_blend_of_constant_tiles_is_constant():
    pipe = FramePipeline(CheapG())
    pl = pipe.plan(1080, 1920, torch.device("cuda"))
    assert pl.ntiles == 60
    tiles = torch.full((60, 1, 256, 256), 0.625, device="cuda")
    got = pipe.blend(tiles, pl)
    assert maxabs(got, torch.full_like(got, 0.625)) <= 1e-6


@pytest.mark.parametrize("n", [2, 1000, 331_776, 2_106_368])
def test_percentiles_match_numpy(n):
    rng = np.random.default_rng(n)
    a = (rng.standard_normal(n) * 0.05 + 0.5).astype(np.float32)
    a[: n // 7] = a[0]  # heavy ties
    pipe = FramePipeline(CheapG())
    for lo, hi in ((0.5, 99.5), (0.1, 99.0), (0.0, 100.0)):
        got = pipe.percentiles(torch.from_numpy(a).cuda(), lo, hi).cpu().numpy()
        ref = np.array([np.percentile(a, lo), np.percentile(a, hi)])
        assert np.abs(got - ref).max() <= 5e-7 * max(1.0, np.abs(ref).max())  # a few fp32 ulps: numpy lerps in fp32
    neg = -a
    got = pipe.percentiles(torch.from_numpy(neg).cuda(), 0.5, 99.5).cpu().numpy()
    assert np.abs(got - np.array([np.percentile(neg, 0.5), np.percentile(neg, 99.5)])).max() <= 5e-7


def test_frame_end_to_end_matches_reference_fixture(net, golden):
    rgb = torch.from_numpy(gi.small_frame())
    pipe = FramePipeline(net)
    pl = pipe.plan(268, 300, torch.device("cuda"))
    gray_p, _ = pipe.normalise_pad(rgb.cuda(), gi.LAMBDA)
    fake_p = pipe.blend(pipe.run_generator(pipe.gather_tiles(gray_p, pl)), pl)
    assert maxabs(fake_p[::2, ::2], golden["frame_fake_s2"][0, 0]) <= 2e-6
    pct = pipe.percentiles(fake_p, 0.5, 99.5).cpu().numpy()
    assert np.abs(pct - golden["frame_percentiles"]).max() <= 5e-7
    col = pipe.tonemap(rgb.cuda(), gi.LAMBDA)
    assert col.shape == (3, 268, 300)
    # stretch divides by (p99.5 - p0.5) ~ 1e-2 of a nearly flat random-init output: errors are amplified ~100x
    assert maxabs(col[:, ::2, ::2], golden["frame_color_s2"]) <= 2e-4
    u8 = pipe.tonemap(rgb.cuda(), gi.LAMBDA, uint8=True)
    assert u8.shape == (268, 300, 3) and u8.dtype == torch.uint8
    assert np.abs(u8.cpu().numpy()[::2, ::2].astype(int) - golden["frame_u8_s2"].astype(int)).max() <= 1


@pytest.mark.parametrize("h,w,padded,ntiles", [(769, 1025, (784, 1040), 24), (1080, 1920, (1088, 1936), 60),
                                               (2160, 3840, (2176, 3856), 220)])
def test_frame_full_sizes_properties(h, w, padded, ntiles):
    """BASELINE.json's full sizes (HDR-Survey 1/4 resolution, 1080p, 4K) through size-independent properties: tile
    count of SURVEY.md section 8d, range of the normalisation, equality of the batched-tile path with tile-at-a-time
    execution (what the reference does), blend as a partition of unity, finite non-negative colour output, and the 8-bit
    stretch reaching both ends of the range."""
    rgb = torch.from_numpy(synth.hdr_frame(h, w, seed=0)).cuda()
    net_bf = UNet(*G_ARGS, up_mode=0, precision="bf16").cuda().eval()
    net_bf.load_state_dict(make_generator_state_dict())
    pipe = FramePipeline(net_bf)
    pl = pipe.plan(h, w, rgb.device)
    assert (pl.h1, pl.w1, pl.ntiles) == (padded[0], padded[1], ntiles)
    gray_p, _ = pipe.normalise_pad(rgb, 50.0)
    assert gray_p.min().item() == 0.0 and abs(gray_p.max().item() - 1.0) < 1e-6
    tiles = pipe.gather_tiles(gray_p, pl)
    batched = pipe.run_generator(tiles)
    probe = [0, ntiles // 3, ntiles - 1]
    single = torch.cat([net_bf.tonemap_tiles(tiles[i:i + 1]) for i in probe])
    assert torch.equal(batched[probe], single)
    ones = pipe.blend(torch.ones_like(batched), pl)
    assert (ones - 1.0).abs().max().item() <= 2e-6
    lo, hi = batched.min().item(), batched.max().item()
    blended = pipe.blend(batched, pl)
    assert lo - 1e-6 <= blended.min().item() and blended.max().item() <= hi + 1e-6     # a convex combination of tile values
    col = pipe.tonemap(rgb, 50.0)
    assert col.shape == (3, h, w) and torch.isfinite(col).all() and col.min().item() >= 0.0
    u8 = pipe.tonemap(rgb, 50.0, uint8=True)
    assert u8.shape == (h, w, 3) and u8.min().item() == 0 and u8.max().item() == 255
    del batched, tiles, blended, col, u8
    torch.cuda.empty_cache()


def test_video_clip_path_matches_oracle():
    """run_model_on_video on a 3-frame 268x300 clip: per-frame normalise, recurrent tile chains, per-frame blend and
    post-process, against the oracle's sequential restatement (5-D tiling, model_save_util.py:488-565)."""
    from uncltmo_b200.generator import UNetVideo
    sd = make_generator_state_dict()
    clip = synth.hdr_clip(3, 268, 300, seed=9)
    lam = 371.4
    vid = UNetVideo(*G_ARGS, up_mode=0, precision="fp32").cuda().eval()
    vid.load_state_dict(sd)
    got = FramePipeline(vid).tonemap_clip(torch.from_numpy(clip).cuda(), lam)
    grays, rgbs = [], []
    for t in range(3):
        rgb, g = oracle.log_lambda_normalise(torch.from_numpy(clip[t]), lam)
        gp, dy, dx = oracle.resize_im(g)
        rgbs.append(oracle.resize_im(rgb)[0])
        grays.append(gp[None, None])
    x5 = torch.cat(grays, 1)  # [1,T,1,H1,W1]
    fakes = oracle.tile_and_blend(x5, lambda tl: oracle.unet_video_forward(sd, tl)[0])
    for t in range(3):
        want = oracle.postprocess_frame(fakes[:, t], rgbs[t], dy, dx)
        assert maxabs(got[t], want) <= 5e-4, t
    indep = FramePipeline(net_fp32()).tonemap(torch.from_numpy(clip[2]).cuda(), lam)
    assert maxabs(got[2], indep) > 1e-4  # the temporal recurrence is live


def net_fp32():
    n = UNet(*G_ARGS, up_mode=0, precision="fp32").cuda().eval()
    n.load_state_dict(make_generator_state_dict())
    return n


def test_streaming_host_api_matches_single_frame_calls(net):
    frames = [torch.from_numpy(synth.hdr_frame(268, 300, seed=20 + i)).pin_memory() for i in range(5)]
    pipe = FramePipeline(net)
    got = pipe.tonemap_host_frames(frames, 50.0)
    torch.cuda.synchronize()
    for f, g in zip(frames, got):
        want = pipe.tonemap(f.cuda(), 50.0, uint8=True).cpu()
        assert torch.equal(g, want)


@pytest.mark.parametrize("h,w", [(769, 1025), (1080, 1920)])
def test_bf16_full_frame_matches_oracle(h, w):
    """The configuration bench.py reports - a whole 1080p frame (and the HDR-Survey 1/4-resolution size) through the
    bf16 tensor-core generator and the GPU frame path - against the fp32 oracle's sequential restatement of
    run_model_on_single_image2 (tile by tile at batch 1, Python cross-fade, np.percentile).
    Tolerances: the blended generator output (before the stretch) carries the per-tile bf16 error (rel-L2 <= 1e-2 is
    BASELINE.json's gate; measured ~3e-4); the post-process divides by (p99.5 - p0.5) of a nearly flat random-init
    output, which amplifies absolute errors ~50-100x (measured: max error = 1.7 % of the output's spread), and the 8-bit
    stretch between the 0.1 / 99.0 percentiles amplifies once more: the colour frame is held to rel-L2 <= 2e-2 (measured
    6.6e-3 / 8.7e-3) and the 8-bit image to +-10 levels with >= 97 % of the pixels within +-2 and >= 85 % within +-1
    (measured: max 7 / 8, 98.6 % / 98.5 %, 89 % / 88 %)."""
    sd = make_generator_state_dict()
    rgb = torch.from_numpy(synth.hdr_frame(h, w, seed=0))
    want = oracle.tonemap_frame(rgb, gi.LAMBDA, lambda t: oracle.unet_forward(sd, t)[0])
    want_u8 = oracle.frame_path.to_uint8_stretch(want)
    net_bf = UNet(*G_ARGS, up_mode=0, precision="bf16").cuda().eval()
    net_bf.load_state_dict(sd)
    pipe = FramePipeline(net_bf)
    # stage 1: blended generator output against the oracle's
    _, g = oracle.log_lambda_normalise(rgb, gi.LAMBDA)
    g_p, _, _ = oracle.resize_im(g)
    fake_want = oracle.tile_and_blend(g_p[None], lambda t: oracle.unet_forward(sd, t)[0])[0, 0]
    pl = pipe.plan(h, w, torch.device("cuda"))
    gray_p, _ = pipe.normalise_pad(rgb.cuda(), gi.LAMBDA)
    fake_got = pipe.blend(pipe.run_generator(pipe.gather_tiles(gray_p, pl)), pl).cpu()
    rel_fake = ((fake_got.double() - fake_want.double()).norm() / fake_want.double().norm()).item()
    # the quantity the stretch amplifies: error relative to the spread of the output, not to its level
    spread = (fake_want.max() - fake_want.min()).item()
    rel_spread = (fake_got - fake_want).abs().max().item() / spread
    col = pipe.tonemap(rgb.cuda(), gi.LAMBDA).cpu()
    rel_col = ((col.double() - want.double()).norm() / want.double().norm()).item()
    u8 = pipe.tonemap(rgb.cuda(), gi.LAMBDA, uint8=True).cpu().numpy()
    d = np.abs(u8.astype(int) - want_u8.astype(int))
    print("bf16 frame %dx%d: blended rel-L2 %.2e, max err / output spread %.2e, colour rel-L2 %.2e, u8 max diff %d, "
          "within +-1: %.4f, +-2: %.4f" % (w, h, rel_fake, rel_spread, rel_col, d.max(), (d <= 1).mean(), (d <= 2).mean()))
    assert rel_fake <= 1e-2
    assert rel_col <= 2e-2
    assert d.max() <= 10 and (d <= 2).mean() >= 0.97 and (d <= 1).mean() >= 0.85


@pytest.mark.parametrize("h,w,negative", [(272, 304, False), (769, 1025, False), (1080, 1920, False), (500, 700, True)])
def test_fused_frame_stages_equal_the_staged_ones(h, w, negative):
    """The three cooperative kernels (normalise -> tiles, blend -> percentiles, post-process -> percentiles -> 8-bit image)
    against the staged calls they replace: the same arithmetic on the same values, so everything is bit-identical -
    tiles, statistics, blended plane, both percentile pairs, the fp32 colour frame and the 8-bit image.  Covers the
    negative-input (shifted statistics) branch and three padded geometries."""
    rgb = torch.from_numpy(synth.hdr_frame(h, w, seed=h + w)).cuda()
    if negative:
        rgb = rgb - 0.37 * rgb.mean()
    pipe = FramePipeline(CheapG())
    pl = pipe.plan(h, w, rgb.device)
    assert pipe.fused_ok(pl)
    gray_p, stats_s = pipe.normalise_pad(rgb, 37.0)
    tiles_s = pipe.gather_tiles(gray_p, pl)
    tiles_f, stats_f = pipe.normalise_tiles(rgb, 37.0, pl)
    assert torch.equal(tiles_s, tiles_f) and torch.equal(stats_s[:3], stats_f[:3])
    out_tiles = pipe.run_generator(tiles_f)
    fake_s = pipe.blend(out_tiles, pl)
    pct_s = pipe.percentiles(fake_s, 0.5, 99.5)
    fake_f, pct_f = pipe.blend_percentiles(out_tiles, pl)
    assert torch.equal(fake_s, fake_f) and torch.equal(pct_s, pct_f)
    col_s = pipe.postprocess(fake_s, rgb, stats_s, pl)
    u8_s = pipe.to_uint8(col_s)
    u8_f, col_f = pipe.post_uint8(fake_f, pct_f, rgb, stats_f, pl, want_col=True)
    assert torch.equal(col_s, col_f) and torch.equal(u8_s, u8_f)
    # whole path, both switches
    full_f = pipe.tonemap(rgb, 37.0, uint8=True)
    pipe.fused = False
    full_s = pipe.tonemap(rgb, 37.0, uint8=True)
    assert torch.equal(full_f, full_s)


def test_frames_batched_through_one_generator_call_equal_single_frames(net):
    """tonemap_frames (several frames' tiles in ONE generator call) and the streaming host API with frames_per_batch = 2
    give exactly the frames tonemap() gives one by one: tiles are independent and every frame keeps its own statistics,
    percentiles and lambda."""
    frames = [torch.from_numpy(synth.hdr_frame(300, 420, seed=s)).cuda() for s in (1, 2, 3)]
    lams = [31.0, 57.0, 44.0]
    pipe = FramePipeline(net)
    single = [pipe.tonemap(f, l, uint8=True) for f, l in zip(frames, lams)]
    batched = pipe.tonemap_frames(frames, lams, uint8=True)
    assert all(torch.equal(a, b) for a, b in zip(single, batched))
    col = pipe.tonemap_frames(frames[:2], lams[0])
    assert torch.equal(col[1], pipe.tonemap(frames[1], lams[0]))
    host = [f.cpu().pin_memory() for f in frames]
    for fpb in (1, 2, 3):
        outs = pipe.tonemap_host_frames(host, lams[0], frames_per_batch=fpb)
        want = [pipe.tonemap(f, lams[0], uint8=True).cpu() for f in frames]
        assert all(torch.equal(a, b) for a, b in zip(outs, want)), fpb
