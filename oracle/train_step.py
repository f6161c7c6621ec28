"""Oracle: TMQI naturalness score and one GAN training step.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference: TMQI.py:210-242 (_StatisticalNaturalness, original=True), GanTrainerImg.py:200-339, 341-408, 452-461.
"""
import numpy as np
import torch
import torch.nn.functional as F

from .discriminator import simple_discriminator_forward
from .generator import unet_forward, unet_video_forward
from .losses import contrastive_d_loss, l1_mean_terms, nce, struct_loss, tv_loss


def tmqi_naturalness(l_ldr):
    """l_ldr: 2-D array in [0,255].  TMQI.py:210-242."""
    from scipy.stats import beta, norm
    phat1, phat2, muhat, sigmahat = 4.4, 10.1, 115.94, 27.99
    l_ldr = np.asarray(l_ldr)
    u = np.mean(l_ldr)
    h, w = l_ldr.shape
    test = np.pad(l_ldr, ((0, 11 - h % 11), (0, 11 - w % 11)), mode="constant")
    view = test.reshape(test.shape[0] // 11, 11, test.shape[1] // 11, 11).transpose(0, 2, 1, 3)
    sig = np.mean(np.std(view, axis=(-1, -2)))
    mode = (phat1 - 1.0) / (phat1 + phat2 - 2.0)
    pc = beta.pdf(sig / 64.29, phat1, phat2) / beta.pdf(mode, phat1, phat2)
    pb = norm.pdf(u, muhat, sigmahat) / norm.pdf(muhat, muhat, sigmahat)
    return pb * pc


def _scores(imgs):
    return [tmqi_naturalness(imgs[i, 0].detach().numpy() * 255) for i in range(imgs.shape[0])]


def info_nce2(fea_fake, fake, k, constant):
    s = _scores(fake)
    pos = fea_fake[s.index(sorted(s)[-1])].unsqueeze(0).repeat(fea_fake.shape[0], 1, 1, 1)
    neg = fea_fake[s.index(sorted(s)[0])].unsqueeze(0).repeat(fea_fake.shape[0], 1, 1, 1)
    return nce(fea_fake, pos, neg, k, constant)


def pseudo_label_loss(fake):
    b, ps = fake.shape[0], fake.shape[-1] // 2
    patches = [fake[i:i + 1, 0:1, j * ps:(j + 1) * ps, k * ps:(k + 1) * ps] for i in range(b) for j in range(2) for k in range(2)]
    s = [tmqi_naturalness(p[0, 0].detach().numpy() * 255) for p in patches]
    label = patches[s.index(sorted(s)[-1])].repeat(len(patches), 1, 1, 1)
    patches = torch.cat(patches, 0)
    lm, lc = l1_mean_terms(patches, label)
    return lm + lc


def g_d_loss(d_fake, d_pos, fea_fake_d, fea_pos_d, fea_neg_d, fea_in_d, fea_fake, fake, ldr_pos, epoch, f=0.1, step1=6, step2=9):
    if epoch <= step2:
        first = epoch <= step1
        err = f * (1.0 if first else 1e-6) * contrastive_d_loss(d_fake, d_pos)
        err = err + f * 0.5 * nce(fea_fake_d, fea_pos_d, fea_in_d, 1, 1e-2)
        err = err + f * 0.5 * 0.2 * nce(fea_fake_d, fea_pos_d, fea_neg_d, 1e3, 2)
        err = err + f * (1e-6 if first else 0.5) * info_nce2(fea_fake, fake, 1, 1e-2)
        lm, lc = l1_mean_terms(fake, ldr_pos)
        err = err + f * (1e-6 if first else 50.0) * lm + f * (1e-6 if first else 1.0) * lc
        err = err + f * 1e-6 * pseudo_label_loss(fake)
    else:
        err = f * 1e-6 * contrastive_d_loss(d_fake, d_pos)
        lm, _ = l1_mean_terms(fake, ldr_pos)
        err = err + f * 50.0 * lm + f * 50.0 * pseudo_label_loss(fake) + f * 0.2 * 1e5 * tv_loss(fake)
    return err


def _generate(gp, hdr, droppath):
    """Image step: hdr [B,1,256,256].  Video step (GanTrainer.py:241-243, 274-276): hdr [B,T,1,256,256] through the
    recurrent generator, outputs flattened to [B*T, ...]."""
    if hdr.dim() == 5:
        fake, fea = unet_video_forward(gp, hdr, droppath)
        return fake.reshape(-1, *fake.shape[2:]), fea.reshape(-1, *fea.shape[2:])
    return unet_forward(gp, hdr, droppath)


def train_step_losses(g_sd, d_sd, hdr, pos, neg, epoch, droppath=None):
    """Loss values and gradients of one D step and one G step at fixed parameters (no optimizer update).
    Returns dict(errD, errG_d, errG_struct, grads_D, grads_G).  A 5-D hdr selects the video trainer's step."""
    gp = {k: v.clone().requires_grad_(k != "gcn.module.0.0.relative_pos") for k, v in g_sd.items()}
    dp = {k: v.clone().requires_grad_(True) for k, v in d_sd.items()}
    d_pos, _ = simple_discriminator_forward(dp, pos)
    with torch.no_grad():
        fake0, _ = _generate(gp, hdr, droppath)
    d_fake, _ = simple_discriminator_forward(dp, fake0)
    err_d = 0.2 * (1.0 if epoch <= 6 else 1e-6) * contrastive_d_loss(d_pos, d_fake)
    grads_d = dict(zip(dp, torch.autograd.grad(err_d, list(dp.values()))))
    fake, fea_fake = _generate(gp, hdr, droppath)
    hdr = hdr.reshape(-1, *hdr.shape[-3:])
    dd = {k: v.detach() for k, v in dp.items()}
    d_fake_bp, fea_fake_d = simple_discriminator_forward(dd, fake)
    d_pos_bp, fea_pos_d = simple_discriminator_forward(dd, pos)
    _, fea_neg_d = simple_discriminator_forward(dd, neg)
    _, fea_in_d = simple_discriminator_forward(dd, hdr)
    err_g = g_d_loss(d_fake_bp, d_pos_bp, fea_fake_d, fea_pos_d, fea_neg_d, fea_in_d, fea_fake, fake, pos, epoch)
    err_s = struct_loss(fake, hdr)
    params = [v for v in gp.values() if v.requires_grad]
    grads_g = dict(zip([k for k, v in gp.items() if v.requires_grad], torch.autograd.grad(err_g + err_s, params)))
    return dict(errD=err_d.item(), errG_d=err_g.item(), errG_struct=err_s.item(), grads_D=grads_d, grads_G=grads_g)
