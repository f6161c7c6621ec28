import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from oracle import train_step as ts
from uncltmo_b200 import synth, losses
from uncltmo_b200.discriminator import SimpleDiscriminator
from uncltmo_b200.struct_loss import StructLoss
from uncltmo_b200.weights import make_discriminator_state_dict, make_generator_state_dict
def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()
g_sd, d_sd = make_generator_state_dict(), make_discriminator_state_dict()
hdr = torch.from_numpy(synth.normalised_batch(2, seed=4)); pos = torch.from_numpy(synth.ldr_batch(2, seed=5)); neg = torch.from_numpy(synth.ldr_batch(2, seed=6))
with torch.no_grad():
    fake0, fea0 = oracle.unet_forward(g_sd, hdr)
D = SimpleDiscriminator(256, 1, 16, "none", "none", 0, 0).cuda(); D.load_state_dict(d_sd)
def terms_oracle(fake, fea):
    d_fake, ff = oracle.simple_discriminator_forward(d_sd, fake)
    d_pos, fp = oracle.simple_discriminator_forward(d_sd, pos)
    _, fn = oracle.simple_discriminator_forward(d_sd, neg)
    _, fi = oracle.simple_discriminator_forward(d_sd, hdr)
    lm, lc = oracle.l1_mean_terms(fake, pos)
    return dict(contrD=oracle.contrastive_d_loss(d_fake, d_pos), nce1=oracle.nce(ff, fp, fi, 1, 1e-2), nce2=oracle.nce(ff, fp, fn, 1e3, 2),
                info2=ts.info_nce2(fea, fake, 1, 1e-2), lmean=lm, lcon=lc, pseudo=ts.pseudo_label_loss(fake), struct=oracle.struct_loss(fake, hdr), tv=oracle.tv_loss(fake))
def terms_cuda(fake, fea):
    d_fake, ff = D(fake)
    with torch.no_grad():
        d_pos, fp = D(pos.cuda()); _, fn = D(neg.cuda()); _, fi = D(hdr.cuda())
    lm, lc = losses.l1_mean_terms(fake, pos.cuda())
    return dict(contrD=losses.contrastive_D_loss(d_fake, d_pos), nce1=losses.infoNCE(ff, fp, fi, None, None, "InfoNCE", 1, 1e-2), nce2=losses.infoNCE(ff, fp, fn, None, None, "InfoNCE", 1e3, 2),
                info2=losses.infoNCE2(fea, fake, None, "InfoNCE", 1, 1e-2), lmean=lm, lcon=lc, pseudo=losses.pseudo_label_loss(fake, None),
                struct=StructLoss([1., 1., 1.])(fake, None, hdr.cuda(), [1., 1., 1.]), tv=losses.L_TV()(fake))
names = list(terms_oracle(fake0, fea0).keys())
for nm in names:
    fr, er = fake0.clone().requires_grad_(True), fea0.clone().requires_grad_(True)
    lo = terms_oracle(fr, er)[nm]; lo.backward()
    fc, ec = fake0.clone().cuda().requires_grad_(True), fea0.clone().cuda().requires_grad_(True)
    lc = terms_cuda(fc, ec)[nm]; lc.backward()
    print("%-8s loss %.6e vs %.6e  dfake rel %.2e  |dfake| %.3e  dfea rel %s" % (nm, lc.item(), lo.item(), rel(fc.grad, fr.grad), fr.grad.norm().item(),
          ("%.2e" % rel(ec.grad, er.grad)) if er.grad is not None and ec.grad is not None else "-"))
