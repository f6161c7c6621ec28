"""Pins the CPU oracle against fixtures produced by the reference itself (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

import golden_inputs as gi
import oracle
from uncltmo_b200.weights import make_generator_state_dict, make_discriminator_state_dict



@pytest.fixture(autouse=True)
def _no_grad():
    """Inference tests run without autograd; tests that need it re-enable it locally."""
    with torch.no_grad():
        yield


def close(a, b, tol=1e-6):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())


def test_generator_image(golden):
    sd = make_generator_state_dict()
    out, up_x, inter = oracle.unet_forward(sd, gi.generator_input(), return_all=True)
    assert close(out, golden["g_img_out"])
    assert close(inter["logit"], golden["g_img_logit"])
    assert close(up_x[:, :, ::8, ::8], golden["g_img_upx_s8"])
    names = ["inc", "down_path_0", "down_path_1", "down_path_2", "down_path_3"]
    acts = dict(zip(names, inter["skips"]))
    acts["gcn"] = inter["gcn"]
    for i in range(4):
        acts["up_path_%d" % i] = inter["ups"][i]
    for name, t in acts.items():
        s = gi.act_stride(t.shape)
        assert close(t[:1, :, ::s, ::s], golden["g_img_act_" + name], 1e-5), name
        sums = golden["g_img_actsum_" + name]
        assert abs(t.double().sum().item() - sums[0]) <= 1e-5 * sums[1], name


def test_generator_video(golden):
    sd = make_generator_state_dict()
    out, feat = oracle.unet_video_forward(sd, gi.video_input())
    assert close(out, golden["g_vid_out"])
    assert close(feat, golden["g_vid_feat"], 1e-5)


def test_relative_pos_matches_weights_module():
    sd = make_generator_state_dict()
    assert torch.equal(oracle.relative_pos_table(), sd["gcn.module.0.0.relative_pos"])


def test_discriminator(golden):
    logit, fea = oracle.simple_discriminator_forward(make_discriminator_state_dict(), gi.ldr_input())
    assert close(logit, golden["d_logit"])
    assert close(fea, golden["d_fea"])


def test_losses(golden):
    sd = make_generator_state_dict()
    x = gi.generator_input()
    fake, _ = oracle.unet_forward(sd, x)
    ld = gi.ldr_input()
    assert close(oracle.struct_loss(fake, x), golden["struct_loss"], 1e-5)
    assert close(oracle.struct_loss(ld[:2], x, (2.0, 4.0, 0.5)), golden["struct_loss_w"], 1e-5)
    a, b = gi.logits_pair()
    assert close(oracle.contrastive_d_loss(a, b), golden["contrastive_d"])
    f1, f2, f3 = gi.nce_features_small()
    assert close(oracle.nce(f1, f2, f3, 1, 1e-2), golden["nce_small_k1"])
    assert close(oracle.nce(f1, f2, f3, 1e3, 2), golden["nce_small_k1e3"])
    g1, g2, g3 = gi.nce_features_map()
    assert close(oracle.nce(g1, g2, g3, 1, 1e-2), golden["nce_map"])
    lm, lc = oracle.l1_mean_terms(fake, ld[:2])
    assert close(lm, golden["l1_mean"])
    assert close(lc, golden["l1_contrast"], 1e-5)
    assert close(oracle.tv_loss(ld), golden["tv"])


def test_frame_path(golden):
    rgb = torch.from_numpy(gi.small_frame())
    rgb2, gray = oracle.log_lambda_normalise(rgb, gi.LAMBDA)
    assert close(gray[:, ::2, ::2], golden["norm_gray_s2"])
    gp, dy, dx = oracle.resize_im(gray)
    rp, _, _ = oracle.resize_im(rgb2)
    assert list(gp.shape) + [dy, dx] == list(golden["pad_shape"])
    assert close(torch.stack([gp[0, 0], gp[0, -1]]), golden["pad_gray_edge"])
    assert close(oracle.tile_and_blend(gp[None], gi.cheap_model)[:, :, ::2, ::2], golden["blend_cheap_s2"])
    big = torch.from_numpy(gi.blend_field(464, 656))
    assert close(oracle.tile_and_blend(big, gi.cheap_model)[:, :, ::4, ::4], golden["blend_cheap_big_s4"])
    big5 = torch.from_numpy(gi.blend_field(272, 400))[:, None].repeat(1, 2, 1, 1, 1)
    big5[:, 1] *= 0.5
    assert close(oracle.tile_and_blend(big5, gi.cheap_model)[..., ::4, ::4], golden["blend_cheap_5d_s4"])
    assert oracle.tile_grid(1088) == ([0, 192, 384, 576, 768], 832)
    assert len(oracle.tile_grid(1936)[0]) + 1 == 10


def test_frame_end_to_end(golden):
    sd = make_generator_state_dict()
    rgb = torch.from_numpy(gi.small_frame())
    rgb2, gray = oracle.log_lambda_normalise(rgb, gi.LAMBDA)
    gp, dy, dx = oracle.resize_im(gray)
    rp, _, _ = oracle.resize_im(rgb2)
    fake = oracle.tile_and_blend(gp[None], lambda t: oracle.unet_forward(sd, t)[0])
    assert close(fake[:, :, ::2, ::2], golden["frame_fake_s2"])
    col = oracle.postprocess_frame(fake, rp, dy, dx)
    assert close(col[:, ::2, ::2], golden["frame_color_s2"], 1e-5)
    u8 = oracle.frame_path.to_uint8_stretch(col)[::2, ::2]
    assert np.abs(u8.astype(int) - golden["frame_u8_s2"].astype(int)).max() <= 1
    assert close(oracle.tonemap_frame(rgb, gi.LAMBDA, lambda t: oracle.unet_forward(sd, t)[0])[:, ::2, ::2],
                 golden["frame_color_s2"], 1e-5)


def test_tmqi_naturalness_and_selection_losses(golden):
    """The score that drives infoNCE2 / pseudo_label_loss (TMQI.py:210-242) and the two losses built on it."""
    from oracle import train_step as ts
    sd = make_generator_state_dict()
    x = gi.generator_input()
    fake, _ = oracle.unet_forward(sd, x)
    ld = gi.ldr_input()
    nat = [oracle.tmqi_naturalness(fake[i, 0].numpy() * 255) for i in range(2)]
    nat += [oracle.tmqi_naturalness(ld[i, 0].numpy() * 255) for i in range(3)]
    q = ld[0, 0].numpy()
    nat += [oracle.tmqi_naturalness(q[j * 128:(j + 1) * 128, k * 128:(k + 1) * 128] * 255) for j in range(2) for k in range(2)]
    assert close(np.array(nat), golden["tmqi_naturalness"], 1e-5)
    assert close(ts.pseudo_label_loss(ld[:2]), golden["pseudo_label_loss"], 1e-5)
    g1, _, _ = gi.nce_features_map()
    assert close(ts.info_nce2(g1[:3], ld, 1, 1e-2), golden["infoNCE2"], 1e-5)
