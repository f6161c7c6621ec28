"""Generate tests/golden/*.npz from the UNMODIFIED reference, imported in place on CPU.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
The reference ships no known-answer vectors, so these fixtures - outputs of the
reference's own classes / functions on seeded synthetic inputs and `uncltmo_b200.weights` parameters -
are what pins the oracle (`tests/test_oracle_golden.py`) and, through it, the CUDA path.

Large tensors are stored strided (`[..., ::s, ::s]`) to keep the fixtures small; inputs are not stored:
tests regenerate them from the same seeds (`tests/golden_inputs.py`).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ref_shims  # noqa: E402
import golden_inputs as gi  # noqa: E402
from uncltmo_b200.weights import make_generator_state_dict, make_discriminator_state_dict  # noqa: E402

torch.set_grad_enabled(False)
R = ref_shims.reference_modules()
torch.Tensor.cuda = lambda self, *a, **k: self  # tiling code hard-codes .cuda() (model_save_util.py:414,418,454)

G_ARGS = dict(n_channels=1, output_dim=1, last_layer="sigmoid", depth=4, layer_factor=4,
              con_operator="square_and_square_root", filters=32, bilinear=0, network="unet", dilation=0,
              to_crop=0, unet_norm="none", stretch_g="none", activation="relu", doubleConvTranspose=1,
              padding_mode="replicate", convtranspose_kernel=2, up_mode=0)


def build_ref_nets():
    sd = make_generator_state_dict()
    g_img = R.gen_img.UNet(**G_ARGS)
    g_img.load_state_dict(sd)
    g_img.eval()
    g_vid = R.gen_vid.UNet(**G_ARGS)
    g_vid.load_state_dict(sd)
    g_vid.eval()
    d = R.disc.SimpleDiscriminator(256, 1, 16, "none", "none", 0, 0)
    d.load_state_dict(make_discriminator_state_dict())
    d.eval()
    return g_img, g_vid, d


def trainer_stub():
    """GanTrainerImg instance without __init__ (needs datasets) - only its loss methods are called."""
    sys.argv = sys.argv[:1]
    import GanTrainerImg
    import GanTrainer
    t = GanTrainerImg.GanTrainer.__new__(GanTrainerImg.GanTrainer)
    return t, GanTrainerImg, GanTrainer


def main():
    out = {}
    g_img, g_vid, d = build_ref_nets()

    # ---- generator, image (Unet_singleFrame.py:177-213) ----
    x = gi.generator_input()
    feats = {}
    hooks = []
    for name in ("inc", "down_path.0", "down_path.1", "down_path.2", "down_path.3", "gcn",
                 "up_path.0", "up_path.1", "up_path.2", "up_path.3", "outc"):
        mod = dict(g_img.named_modules())[name]
        hooks.append(mod.register_forward_hook(lambda m, i, o, name=name: feats.__setitem__(name, o)))
    y, up_x = g_img(x)
    for h in hooks:
        h.remove()
    out["g_img_out"] = y.numpy()
    out["g_img_logit"] = feats["outc"].numpy()
    out["g_img_upx_s8"] = up_x[:, :, ::8, ::8].numpy()
    for name, t in feats.items():
        if name != "outc":
            s = gi.act_stride(t.shape)
            out["g_img_act_" + name.replace(".", "_")] = t[:1, :, ::s, ::s].numpy()
            out["g_img_actsum_" + name.replace(".", "_")] = np.array([t.double().sum().item(), t.double().abs().sum().item()])

    # ---- generator, video (Unet.py:213-289) ----
    xv = gi.video_input()
    yv, fv = g_vid(xv)
    out["g_vid_out"] = yv.numpy()
    out["g_vid_feat"] = fv.numpy()

    # ---- discriminator (Discriminator.py:119-126) ----
    ld = gi.ldr_input()
    logit, fea = d(ld)
    out["d_logit"] = logit.numpy()
    out["d_fea"] = fea.numpy()

    # ---- losses ----
    t, GTI, GT = trainer_stub()
    fake = y
    hdr = x
    sl = R.struct_loss.StructLoss(pyramid_weight_list=torch.tensor([1.0, 1.0, 1.0]), window_size=5)
    out["struct_loss"] = np.array(sl(fake, None, hdr, torch.tensor([1.0, 1.0, 1.0])).item())
    out["struct_loss_w"] = np.array(sl(ld[:2], None, hdr, torch.tensor([2.0, 4.0, 0.5])).item())
    a, b = gi.logits_pair()
    out["contrastive_d"] = np.array(t.contrastive_D_loss(a, b).item())
    f1, f2, f3 = gi.nce_features_small()
    out["nce_small_k1"] = np.array(t.nce(f1, [f2], [f3], "InfoNCE", 1, 1e-2).item())
    out["nce_small_k1e3"] = np.array(t.nce(f1, [f2], [f3], "InfoNCE", 1e3, 2).item())
    g1, g2, g3 = gi.nce_features_map()
    out["nce_map"] = np.array(t.nce(g1, [g2], [g3], "InfoNCE", 1, 1e-2).item())
    l1 = torch.nn.L1Loss()
    ce = GTI.ContrastExtracter()
    out["l1_mean"] = np.array(l1(fake.mean(dim=[-1, -2]), ld[:2].mean(dim=[-1, -2])).item())
    out["l1_contrast"] = np.array(l1(ce(fake).mean(dim=[-1, -2]), ce(ld[:2]).mean(dim=[-1, -2])).item())
    out["tv"] = np.array(GT.L_TV()(ld).item())
    # TMQI statistical naturalness, the score behind infoNCE2 / pseudo_label_loss (TMQI.py:210-242)
    from TMQI import TMQI
    tm = TMQI()
    nat = [tm(hdr[i, 0].numpy().astype(np.float64), (fake[i, 0].numpy() * 255).astype(np.float64))[2] for i in range(2)]
    nat += [tm(ld[i, 0].numpy().astype(np.float64) + 1e-3, (ld[i, 0].numpy() * 255).astype(np.float64))[2] for i in range(3)]
    q = ld[0, 0].numpy()
    nat += [tm(q[j * 128:(j + 1) * 128, k * 128:(k + 1) * 128].astype(np.float64) + 1e-3,
               (q[j * 128:(j + 1) * 128, k * 128:(k + 1) * 128] * 255).astype(np.float64))[2] for j in range(2) for k in range(2)]
    out["tmqi_naturalness"] = np.array(nat)
    out["pseudo_label_loss"] = np.array(t.pseudo_label_loss(ld[:2], ld[:2] + 1e-3).item())
    out["infoNCE2"] = np.array(t.infoNCE2(g1[:3], ld, ld + 1e-3, "InfoNCE", 1, 1e-2).item())

    # ---- frame path: normalise, pad, tile+blend, post-process ----
    rgb = torch.from_numpy(gi.small_frame())
    gray = R.hdr_util.to_gray_tensor(rgb)
    gray = gray - gray.min()
    f = gi.LAMBDA * 255 * 0.1
    gray = torch.log10((gray / gray.max()) * f + 1)
    gray = gray / gray.max()  # model_save_util.py:236-239
    out["norm_gray_s2"] = gray[:, ::2, ::2].numpy()
    rgb_p, dy, dx = R.dl_util.resize_im(rgb, False, 0)
    gray_p, _, _ = R.dl_util.resize_im(gray, False, 0)
    out["pad_shape"] = np.array(list(gray_p.shape) + [dy, dx])
    out["pad_gray_edge"] = np.stack([gray_p[0, 0, :].numpy(), gray_p[0, -1, :].numpy()])

    class Cheap(torch.nn.Module):  # stand-in generator: pins the tiling/blend arithmetic alone
        def forward(self, t, apply_crop=True, diffY=0, diffX=0):
            return gi.cheap_model(t), None

    blended = R.save_util.test_big_size_image2(gray_p.unsqueeze(0), Cheap(), False, dy, dx)
    out["blend_cheap_s2"] = blended[:, :, ::2, ::2].numpy()
    big = torch.from_numpy(gi.blend_field(464, 656))
    out["blend_cheap_big_s4"] = R.save_util.test_big_size_image2(big, Cheap(), False, 0, 0)[:, :, ::4, ::4].numpy()
    big5 = torch.from_numpy(gi.blend_field(272, 400))[:, None].repeat(1, 2, 1, 1, 1)
    big5[:, 1] *= 0.5
    out["blend_cheap_5d_s4"] = R.save_util.test_big_size_image(big5, Cheap(), False, 0, 0)[..., ::4, ::4].numpy()

    # full image entry path on the small frame with the real generator (4 tiles + post-process)
    fake_full = R.save_util.test_big_size_image2(gray_p.unsqueeze(0), g_img, False, dy, dx)
    out["frame_fake_s2"] = fake_full[:, :, ::2, ::2].numpy()
    max_p = np.percentile(fake_full.numpy(), 99.5)
    min_p = np.percentile(fake_full.numpy(), 0.5)
    out["frame_percentiles"] = np.array([min_p, max_p], dtype=np.float64)
    f2 = fake_full.clamp(min_p, max_p)
    st = (f2 - f2.min()) / (f2.max() - f2.min())
    col = R.hdr_util.back_to_color_tensor(rgb_p, st[0], torch.device("cpu"))
    im_max = col.max()
    col = col[:, dy // 2:-(dy - dy // 2), dx // 2:-(dx - dx // 2)].clamp(min=0, max=im_max)
    out["frame_color_s2"] = col[:, ::2, ::2].numpy()
    t01 = col.clamp(0, 1).permute(1, 2, 0).numpy()
    out["frame_u8_s2"] = (R.hdr_util.to_0_1_range_outlier(np.squeeze(t01)) * 255).astype("uint8")[::2, ::2]

    path = os.path.join(HERE, "reference_outputs.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB;", len(out), "arrays")


if __name__ == "__main__":
    main()
