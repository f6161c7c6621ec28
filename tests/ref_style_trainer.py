"""TEST INFRASTRUCTURE: a trainer that makes the calls the reference trainers make, in the order they make them.

/root/reference does not exist on the GPU box, so the reference's own `GanTrainer` object cannot be instantiated over
the drop-in modules there.  This class restates `train_D` / `D_real_fake_pass` / `train_G` / `update_g_d_loss` /
`update_struct_loss` call for call (GanTrainerImg.py:200-217, 231-260, 262-292, 302-339, 452-461; the video trainer
GanTrainer.py:233-300 differs only in handing the 5-D clip batch to netG and flattening its outputs) over whatever
`netG`, `netD`, `struct_loss` objects it is given - including the reference's quirks that matter to a drop-in module:
weights kept as tensors (`adv_weight_list[0].float()`), `.float()` / `.reshape` on inputs, `netD(fake.detach())`,
`D(real_neg)` computed and dropped, `errG_d.backward(retain_graph=True)` followed by a second backward through the
same generator graph for the structural loss, and the loss glue written with stock torch ops on the modules' outputs
(host-side TMQI on `.detach().cpu().numpy()` copies).  With the reference's own modules on the CPU it reproduces the
fixtures of tests/golden/make_golden_train.py (tests/test_oracle_golden.py); with the uncltmo_b200 modules on the GPU
it is the "swap four imports" claim of INTEGRATION.md section 1.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

import oracle
from oracle.discriminator import contrast_map


class RefStyleTrainer:
    def __init__(self, netG, netD, struct_loss, optimizerG, optimizerD, video=False, losses=None):
        self.netG, self.netD, self.struct_loss = netG, netD, struct_loss
        self.optimizerG, self.optimizerD = optimizerG, optimizerD
        self.video = video
        self.loss_g_d_factor, self.struct_loss_factor = 0.1, 1.0
        self.epoch_step1, self.epoch_step2 = 6, 9
        self.adv_weight_list = torch.tensor([0.2, 0.2, 0.2])
        self.pyramid_weight_list = torch.tensor([1.0, 1.0, 1.0])
        self.pre_train_mode, self.train_with_D, self.manual_d_training = False, True, False
        self.final_shape_addition, self.to_crop = 0, 0
        self.G_loss_d, self.G_loss_struct, self.D_losses = [], [], []
        self.losses = losses   # None: torch glue below; else a module with the reference's loss-method names

    # ---- GanTrainerImg.py:219-229
    def contrastive_D_loss(self, real_logits, fake_logits):
        if self.losses is not None:
            return self.losses.contrastive_D_loss(real_logits, fake_logits)
        r, f = real_logits.reshape(-1), fake_logits.reshape(-1)

        def loss_half(t1, t2):
            t = torch.cat((t1[:, None], t2[None, :].repeat(t1.shape[0], 1)), dim=-1)
            return F.cross_entropy(t, torch.zeros(t1.shape[0], device=t.device, dtype=torch.long))

        return loss_half(r, f) + loss_half(-f, -r)

    # ---- GanTrainerImg.py:410-439
    def nce(self, fea_anchor, feas_positive, feas_negative, cl_loss_type, k, constant):
        if self.losses is not None:
            return self.losses.nce(fea_anchor, feas_positive, feas_negative, cl_loss_type, k, constant)
        b = fea_anchor.shape[0]
        c = constant
        neg = [torch.sum((fea_anchor * f) * (1 / (c + k * torch.abs(fea_anchor - f))), dim=1).mean(dim=[-1, -2]).unsqueeze(1)
               for f in feas_negative]
        loss = 0
        for f in feas_positive:
            pos = [torch.sum((fea_anchor * f) * (1 / (c + k * torch.abs(fea_anchor - f))), dim=1).mean(dim=[-1, -2]).unsqueeze(1)]
            logits = torch.cat(pos + neg, dim=1)
            loss += F.cross_entropy(logits, torch.zeros(b, device=logits.device, dtype=torch.long))
        return loss / len(feas_positive)

    def infoNCE(self, fea_fake, fea_real, fea_neg, fake, hdr_input, cl_loss_type, k, constant):
        return self.nce(fea_fake, [fea_real], [fea_neg], cl_loss_type, k, constant)

    # ---- GanTrainerImg.py:384-408 (host TMQI on numpy copies, as the reference does)
    def infoNCE2(self, fea_fake, fake, hdr_input, cl_loss_type, k, constant):
        if self.losses is not None:
            return self.losses.infoNCE2(fea_fake, fake, hdr_input, cl_loss_type, k, constant)
        fake_imgs = fake.permute(0, 2, 3, 1).detach().cpu().numpy()
        s = [oracle.tmqi_naturalness(fake_imgs[i, :, :, 0] * 255) for i in range(fake.shape[0])]
        ss = sorted(s)
        pos = fea_fake[s.index(ss[-1]), :, :, :].unsqueeze(0).repeat(fea_fake.shape[0], 1, 1, 1)
        neg = fea_fake[s.index(ss[0]), :, :, :].unsqueeze(0).repeat(fea_fake.shape[0], 1, 1, 1)
        return self.nce(fea_fake, [pos], [neg], cl_loss_type, k, constant)

    @staticmethod
    def _contrast(x):   # ContrastExtracter, GanTrainerImg.py:24-56 (C = 1)
        return contrast_map(x)

    # ---- GanTrainerImg.py:341-368
    def pseudo_label_loss(self, fake, hdr_input):
        if self.losses is not None:
            return self.losses.pseudo_label_loss(fake, hdr_input)
        fake_imgs = fake.permute(0, 2, 3, 1).detach().cpu().numpy()
        ps = 128
        patches, scores = [], []
        for i in range(fake.shape[0]):
            for j in range(2):
                for k in range(2):
                    scores.append(oracle.tmqi_naturalness(fake_imgs[i, j * ps:(j + 1) * ps, k * ps:(k + 1) * ps, 0] * 255))
                    patches.append(fake[i:i + 1, 0:1, j * ps:(j + 1) * ps, k * ps:(k + 1) * ps])
        label = patches[scores.index(sorted(scores)[-1])].repeat(len(patches), 1, 1, 1)
        patches = torch.cat(patches, 0)
        l1 = nn.L1Loss()
        loss = l1(patches.mean(dim=[-1, -2]), label.mean(dim=[-1, -2]))
        loss += l1(self._contrast(patches).mean(dim=[-1, -2]), self._contrast(label).mean(dim=[-1, -2]))
        return loss

    def tv(self, x):   # GanTrainer.py:669-682
        if self.losses is not None:
            return self.losses.L_TV()(x)
        return oracle.tv_loss(x)

    # ---- GanTrainerImg.py:200-217, 231-260
    def train_D(self, hdr_input, real_ldr_pos, real_ldr_neg, epoch):
        self.netD.zero_grad()
        self.D_real_fake_pass(real_ldr_pos.float(), real_ldr_neg.float(), hdr_input.float(), epoch)
        self.optimizerD.step()
        self.D_losses.append(self.errD.item())

    def _flat(self, t):
        return t.reshape(-1, t.shape[2], t.shape[3], t.shape[4])

    def _G(self, hdr_input):
        if self.video:   # GanTrainer.py:241-243, 274-276
            fake, fea = self.netG(hdr_input, diffY=self.final_shape_addition, diffX=self.final_shape_addition)
            return self._flat(fake), self._flat(fea)
        return self.netG(self._flat(hdr_input), diffY=self.final_shape_addition, diffX=self.final_shape_addition)

    def D_real_fake_pass(self, real_ldr_pos, real_ldr_neg, hdr_input, epoch):
        d_real_pos, _ = self.netD(self._flat(real_ldr_pos))
        d_real_neg, _ = self.netD(self._flat(real_ldr_neg))   # computed and never used
        fake, _ = self._G(hdr_input)
        d_fake, _ = self.netD(fake.detach())
        w = self.adv_weight_list[0].float() * (1.0 if epoch <= self.epoch_step1 else 1e-6)
        self.errD = w * self.contrastive_D_loss(d_real_pos, d_fake)
        self.errD.backward()

    # ---- GanTrainerImg.py:262-292
    def train_G(self, hdr_input, hdr_original_gray_norm, real_ldr_pos, real_ldr_neg, epoch):
        self.netG.zero_grad()
        fake, fea_fake = self._G(hdr_input.float())
        d_fake_bp, d_fea_fake = self.netD(fake.float())
        d_real_pos_bp, d_fea_real_pos = self.netD(self._flat(real_ldr_pos))
        d_real_neg_bp, d_fea_real_neg = self.netD(self._flat(real_ldr_neg))
        _, d_fea_input = self.netD(self._flat(hdr_input.float()))
        self.update_g_d_loss(d_fake_bp, d_real_pos_bp, d_real_neg_bp, d_fea_fake, d_fea_real_pos, d_fea_real_neg, d_fea_input,
                             fea_fake, fake, self._flat(hdr_input.float()), self._flat(real_ldr_pos), self._flat(real_ldr_neg),
                             epoch)
        hdr_input = self._flat(hdr_input)
        self.update_struct_loss(hdr_input, self._flat(hdr_original_gray_norm), fake)
        self.optimizerG.step()

    # ---- GanTrainerImg.py:302-339
    def update_g_d_loss(self, d_fake_bp, d_real_pos_bp, d_real_neg_bp, d_fea_fake, d_fea_real_pos, d_fea_real_neg, d_fea_input,
                        fea_fake, fake, hdr_input, ldr_pos, ldr_neg, epoch):
        f = self.loss_g_d_factor
        l1_loss = nn.L1Loss()
        if epoch <= self.epoch_step1:
            self.errG_d = f * self.contrastive_D_loss(d_fake_bp, d_real_pos_bp)
            self.errG_d += f * 0.5 * self.infoNCE(d_fea_fake, d_fea_real_pos, d_fea_input, fake, hdr_input, cl_loss_type='InfoNCE', k=1, constant=1e-2)
            self.errG_d += f * 0.5 * (0.2 * self.infoNCE(d_fea_fake, d_fea_real_pos, d_fea_real_neg, fake, hdr_input, cl_loss_type='InfoNCE', k=1e3, constant=2))
            self.errG_d += f * 1e-6 * self.infoNCE2(fea_fake, fake, hdr_input, cl_loss_type='InfoNCE', k=1, constant=1e-2)
            self.errG_d += f * 1e-6 * l1_loss(fake.mean(dim=[-1, -2]), ldr_pos.mean(dim=[-1, -2]))
            self.errG_d += f * 1e-6 * l1_loss(self._contrast(fake).mean(dim=[-1, -2]), self._contrast(ldr_pos).mean(dim=[-1, -2]))
            self.errG_d += f * 1e-6 * self.pseudo_label_loss(fake, hdr_input)
        elif epoch <= self.epoch_step2:
            self.errG_d = f * 1e-6 * self.contrastive_D_loss(d_fake_bp, d_real_pos_bp)
            self.errG_d += f * 0.5 * self.infoNCE(d_fea_fake, d_fea_real_pos, d_fea_input, fake, hdr_input, cl_loss_type='InfoNCE', k=1, constant=1e-2)
            self.errG_d += f * 0.5 * (0.2 * self.infoNCE(d_fea_fake, d_fea_real_pos, d_fea_real_neg, fake, hdr_input, cl_loss_type='InfoNCE', k=1e3, constant=2))
            self.errG_d += f * 0.1 * (5 * self.infoNCE2(fea_fake, fake, hdr_input, cl_loss_type='InfoNCE', k=1, constant=1e-2))
            self.errG_d += f * 0.5 * (1e2 * l1_loss(fake.mean(dim=[-1, -2]), ldr_pos.mean(dim=[-1, -2])))
            self.errG_d += f * 0.5 * (2 * l1_loss(self._contrast(fake).mean(dim=[-1, -2]), self._contrast(ldr_pos).mean(dim=[-1, -2])))
            self.errG_d += f * 1e-6 * self.pseudo_label_loss(fake, hdr_input)
        else:
            self.errG_d = f * 1e-6 * self.contrastive_D_loss(d_fake_bp, d_real_pos_bp)
            self.errG_d += f * 0.5 * (1e2 * l1_loss(fake.mean(dim=[-1, -2]), ldr_pos.mean(dim=[-1, -2])))
            self.errG_d += f * 0.5 * (1e2 * self.pseudo_label_loss(fake, hdr_input))
            self.errG_d += f * 0.2 * (1e5 * self.tv(fake))
        retain_graph = bool(self.struct_loss_factor)
        self.errG_d.backward(retain_graph=retain_graph)
        self.G_loss_d.append(self.errG_d.item())

    # ---- GanTrainerImg.py:452-461
    def update_struct_loss(self, hdr_input, hdr_input_original_gray_norm, fake):
        if self.struct_loss_factor:
            self.errG_struct = self.struct_loss_factor * self.struct_loss(fake, hdr_input_original_gray_norm, hdr_input,
                                                                          self.pyramid_weight_list)
            self.errG_struct.backward()
            self.G_loss_struct.append(self.errG_struct.item())
