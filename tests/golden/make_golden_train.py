"""Training-step fixtures from the UNMODIFIED reference trainers, imported in place on CPU.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_train.py

What is stored (tests/golden/reference_train_outputs.npz):
  * `ref/<trainer>/e<epoch>/b<B>/{errD,errG_d,errG_struct}`: the values the reference's OWN `train_D` + `train_G`
    leave in `self.errD / self.errG_d / self.errG_struct` (GanTrainerImg.py:200-217, 262-292; GanTrainer.py:233-300),
    fp32, on seeded synthetic batches and `uncltmo_b200.weights` parameters.  The trainer object is created with
    `__new__` + attribute injection (its `__init__` needs datasets); the reference modules,
    `StructLoss` and loss methods are the reference's.  Optimizers are SGD(lr=0) so both halves of the step see the
    same parameters (the fixtures pin the loss schedule and the gradients, not Adam).  DropPath p is set to 0 on the
    module instances (two forwards per step draw different masks otherwise).
  * `ref/.../gD/<name>`, `ref/.../gG/<name>`: [L2 norm, sum, sum of |.|] of every parameter gradient the reference
    accumulated (`p.grad` after the step) and a strided sample of its values.
  * `o64/.../*`: the same quantities from the repo's oracle (`oracle.train_step_losses`) evaluated in float64 -
    the anchor the GPU tests compare with at 16 images without paying ~40 s of host time per case on the GPU box.
  * `obf/img/e<epoch>/b16/gG/<name>`: oracle gradients with every 3x3 / k2s2 / 1x1 tensor-core operand rounded to
    bf16 (the rounding points of the mixed-precision CUDA path, `oracle.generator.bf16_operands`), float64 accumulate.

The image trainer's `epoch > epoch_step2` branch references `L_TV`, which GanTrainerImg.py never defines
(SURVEY.md R10): the name is injected into the module namespace from GanTrainer.py:669-682, where the same author
defines it.  Nothing under /root/reference is modified.
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

R = ref_shims.reference_modules()
sys.argv = sys.argv[:1]
import GanTrainerImg  # noqa: E402
import GanTrainer  # noqa: E402

if not hasattr(GanTrainerImg, "L_TV"):
    GanTrainerImg.L_TV = GanTrainer.L_TV   # see the module docstring

G_ARGS = dict(n_channels=1, output_dim=1, last_layer="sigmoid", depth=4, layer_factor=4,
              con_operator="square_and_square_root", filters=32, bilinear=0, network="unet", dilation=0,
              to_crop=0, unet_norm="none", stretch_g="none", activation="relu", doubleConvTranspose=1,
              padding_mode="replicate", convtranspose_kernel=2, up_mode=0)


def make_trainer(video):
    mod = GanTrainer if video else GanTrainerImg
    t = mod.GanTrainer.__new__(mod.GanTrainer)
    net_cls = R.gen_vid.UNet if video else R.gen_img.UNet
    t.netG = net_cls(**G_ARGS)
    t.netG.load_state_dict(make_generator_state_dict())
    t.netG.train()
    for m in t.netG.modules():   # DropPath(0.05) on both residual branches of the graph block -> identity
        if hasattr(m, "drop_prob"):
            m.drop_prob = 0.0
    t.netD = R.disc.SimpleDiscriminator(256, 1, 16, "none", "none", 0, 0)
    t.netD.load_state_dict(make_discriminator_state_dict())
    t.netD.train()
    t.optimizerG = torch.optim.SGD([p for p in t.netG.parameters() if p.requires_grad], lr=0.0)
    t.optimizerD = torch.optim.SGD(t.netD.parameters(), lr=0.0)
    # run_imageTMO_train.sh / main_train.py defaults
    t.loss_g_d_factor, t.struct_loss_factor = 0.1, 1.0
    t.epoch_step1, t.epoch_step2 = 6, 9
    t.adv_weight_list = torch.tensor([0.2, 0.2, 0.2])
    t.pyramid_weight_list = torch.tensor([1.0, 1.0, 1.0])
    t.struct_loss = R.struct_loss.StructLoss(pyramid_weight_list=t.pyramid_weight_list, window_size=5)
    t.pre_train_mode, t.train_with_D, t.manual_d_training = False, True, False
    t.final_shape_addition, t.to_crop = 0, 0
    t.G_loss_d, t.G_loss_struct, t.D_losses = [], [], []
    return t


def stats(g):
    g = g.detach().double().reshape(-1)
    step = max(1, g.numel() // 64)   # tests/test_oracle_train_golden.py:_stats samples the same elements
    return np.concatenate([[g.norm().item(), g.sum().item(), g.abs().sum().item()], g[::step][:64].numpy()])


def run_reference(video, epoch, b):
    hdr, pos, neg = gi.train_batch(b, video)
    t = make_trainer(video)
    t.train_D(hdr, pos, neg, epoch)
    out = {"errD": np.array(t.errD.item())}
    for k, p in t.netD.named_parameters():
        out["gD/" + k] = stats(p.grad)
    t.train_G(hdr, hdr, pos, neg, epoch)
    out["errG_d"] = np.array(t.errG_d.item())
    out["errG_struct"] = np.array(t.errG_struct.item())
    for k, p in t.netG.named_parameters():
        if p.grad is not None:
            out["gG/" + k] = stats(p.grad)
    return out


def run_oracle64(video, epoch, b, bf16=False, droppath=False):
    import oracle
    hdr, pos, neg = gi.train_batch(b, video)
    dp = [m.double() for m in gi.droppath_masks(b)] if droppath else None
    g_sd = {k: v.double() for k, v in make_generator_state_dict().items()}
    d_sd = {k: v.double() for k, v in make_discriminator_state_dict().items()}
    h = hdr.double() if video else hdr.reshape(-1, 1, 256, 256).double()
    with oracle.bf16_operands(bf16):
        r = oracle.train_step_losses(g_sd, d_sd, h, pos.reshape(-1, 1, 256, 256).double(),
                                     neg.reshape(-1, 1, 256, 256).double(), epoch, droppath=dp)
    out = {k: np.array(r[k]) for k in ("errD", "errG_d", "errG_struct")}
    for k, g in r["grads_D"].items():
        out["gD/" + k] = stats(g)
    for k, g in r["grads_G"].items():
        out["gG/" + k] = stats(g)
    return out


def main():
    import time
    out = {}
    cases = [(False, e, b) for e in (0, 7, 10) for b in (4, 16)] + [(True, e, b) for e in (0, 7, 10) for b in (4, 16)]
    only = os.environ.get("GOLDEN_TRAIN_ONLY")
    for video, epoch, b in cases:
        tag = "%s/e%d/b%d" % ("vid" if video else "img", epoch, b)
        if only and only not in tag:
            continue
        t0 = time.time()
        for k, v in run_reference(video, epoch, b).items():
            out["ref/%s/%s" % (tag, k)] = v
        t1 = time.time()
        for k, v in run_oracle64(video, epoch, b).items():
            out["o64/%s/%s" % (tag, k)] = v
        t2 = time.time()
        if b == 16 and not video:
            for k, v in run_oracle64(video, epoch, b, bf16=True).items():
                out["obf/%s/%s" % (tag, k)] = v
        print(tag, "reference %.0f s, oracle64 %.0f s, bf16-operand oracle %.0f s" % (t1 - t0, t2 - t1, time.time() - t2),
              {k: float(out["ref/%s/%s" % (tag, k)]) for k in ("errD", "errG_d", "errG_struct")},
              {k: float(out["o64/%s/%s" % (tag, k)]) for k in ("errD", "errG_d", "errG_struct")}, flush=True)
    # train-mode DropPath with injected per-sample masks (the reference draws them from the torch RNG inside timm's
    # DropPath; the fixtures above run with p = 0): oracle only, float64 and bf16-operand
    tag = "img/e0/b16dp"
    if not only or only in tag:
        for k, v in run_oracle64(False, 0, 16, droppath=True).items():
            out["o64/%s/%s" % (tag, k)] = v
        for k, v in run_oracle64(False, 0, 16, bf16=True, droppath=True).items():
            out["obf/%s/%s" % (tag, k)] = v
        print(tag, {k: float(out["o64/%s/%s" % (tag, k)]) for k in ("errD", "errG_d", "errG_struct")}, flush=True)
    path = os.path.join(HERE, "reference_train_outputs.npz")
    if only and os.path.exists(path):
        old = dict(np.load(path))
        old.update(out)
        out = old
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB;", len(out), "arrays")


if __name__ == "__main__":
    main()
