"""Loss methods of GanTrainer / GanTrainerImg with the reference call signatures, on sm_100a kernels (fwd + bwd).

Reference: GanTrainerImg.py:219-229 (contrastive_D_loss), :410-439 (nce), :370-382 (infoNCE), :384-408 (infoNCE2),
:341-368 (pseudo_label_loss), :308-313 (mean / contrast L1 terms), GanTrainer.py:669-682 (L_TV).
The reference picks the positives / negatives of infoNCE2 and the pseudo label with a HOST numpy TMQI call per image
(80 calls and a device->host copy per step); here the naturalness score is a kernel and the arg-max stays on the device.
"""
import torch

from . import dist as udist
from .autograd_losses import (ContrastiveDFn, L1MeanFn, NceFn, NceSelfFn, PlaneMeanContrastFn, TVFn, tmqi_naturalness)


def contrastive_D_loss(real_logits, fake_logits):
    return ContrastiveDFn.apply(real_logits, fake_logits)


def nce(fea_anchor, feas_positive, feas_negative, cl_loss_type, k, constant):
    if cl_loss_type != "InfoNCE" or len(feas_positive) != 1 or len(feas_negative) != 1:
        raise NotImplementedError("only the shipped call pattern (InfoNCE, one positive, one negative) is built")
    return NceFn.apply(fea_anchor, feas_positive[0], feas_negative[0], k, constant)


def infoNCE(fea_fake, fea_real, fea_neg, fake, hdr_input, cl_loss_type, k, constant):
    return nce(fea_fake, [fea_real], [fea_neg], cl_loss_type, k, constant)


def nce_from_indices(fea_fake, pos_index, neg_index, cl_loss_type, k, constant):
    """infoNCE2 after the selection: positive / negative = one sample of the batch, broadcast (GanTrainerImg.py:400-405).
    pos_index / neg_index: python ints or 0-dim / 1-element device tensors (no host sync in the latter case)."""
    def pick(i):
        if torch.is_tensor(i):
            return fea_fake.index_select(0, i.reshape(1))
        return fea_fake[i:i + 1]
    return nce(fea_fake, [pick(pos_index)], [pick(neg_index)], cl_loss_type, k, constant)


def infoNCE2(fea_fake, fake, hdr_input, cl_loss_type, k, constant):
    """GanTrainerImg.py:384-408: positive / negative = the batch samples with the highest / lowest TMQI naturalness."""
    n = tmqi_naturalness(fake)
    blocked = fea_fake.dim() == 5 and fea_fake.dtype == torch.bfloat16
    if blocked and cl_loss_type != "InfoNCE":
        raise NotImplementedError("only InfoNCE is built")
    if udist.is_parallel():
        # data parallel: arg-max / arg-min over the GLOBAL batch (the reference selects over the whole batch under
        # nn.DataParallel); the two chosen feature maps are broadcast from the ranks that hold them and the gradient that
        # reaches them is returned to those ranks.  The value is this rank's share of the global-batch mean.
        g = udist.all_gather_flat(n)
        g_idx = torch.stack([torch.argmax(g), torch.argmin(g)])
        rows = udist.BroadcastRowsFn.apply(fea_fake, g_idx)
        b = fea_fake.shape[0]
        if blocked:
            sel = torch.arange(b, b + 2, device=fea_fake.device, dtype=torch.int64)   # (no H2D copy: graph-capture safe)
            return NceSelfFn.apply(fea_fake, sel, fea_fake.shape[2] * fea_fake.shape[3], k, constant, rows)
        return nce(fea_fake, [rows[0:1]], [rows[1:2]], cl_loss_type, k, constant)
    if blocked:
        # the bf16 training path hands over up_x as the C8-blocked tensor it was computed in (UNet.forward_blocked)
        sel = torch.stack([torch.argmax(n), torch.argmin(n)])
        return NceSelfFn.apply(fea_fake, sel, fea_fake.shape[2] * fea_fake.shape[3], k, constant)
    return nce_from_indices(fea_fake, torch.argmax(n), torch.argmin(n), cl_loss_type, k, constant)


def plane_mean_contrast(x):
    return PlaneMeanContrastFn.apply(x)


def l1_mean_terms(fake, ldr):
    """(L1 of per-image means, L1 of per-image mean local variance): GanTrainerImg.py:308-313."""
    fm, fc = PlaneMeanContrastFn.apply(fake)
    lm, lc = PlaneMeanContrastFn.apply(ldr)
    return L1MeanFn.apply(fm, lm), L1MeanFn.apply(fc, lc)


def pseudo_label_loss(fake, hdr_input):
    """GanTrainerImg.py:341-368: split every fake into 2x2 quadrants, take the quadrant with the best naturalness as
    the pseudo label, L1 between all quadrants and the label in mean and in mean local variance."""
    b = fake.shape[0]
    ps = fake.shape[-1] // 2
    patches = fake.reshape(b, 1, 2, ps, 2, ps).permute(0, 2, 4, 1, 3, 5).reshape(b * 4, 1, ps, ps).contiguous()
    scores = tmqi_naturalness(patches)
    pm, pc = PlaneMeanContrastFn.apply(patches)
    if udist.is_parallel():
        # Data parallel: the pseudo label is the best quadrant of the GLOBAL batch (the reference selects over the whole
        # batch under nn.DataParallel).  The loss needs two statistics per quadrant, so those are gathered:
        #   L = mean_i |pm_i - lm| + mean_i |pc_i - lc|,  (lm, lc) = statistics of the globally best quadrant.
        # Rank r returns A_r + (B - B.detach()):  A_r = mean over ITS quadrants with the label held constant (value and
        # gradient of its own terms, same "local mean" convention as every other per-sample term: the trainer scales by
        # 1/world and the gradient all-reduce sums);  B = (1/world-scaled) ALL terms with only the label live - zero in
        # value, it gives the rank that owns the label the label-path gradient of every rank's terms.
        world = udist.dist.get_world_size()
        best = torch.argmax(udist.all_gather_flat(scores)).reshape(1)
        gm, gc = udist.all_gather_cat(pm.reshape(-1, 1)), udist.all_gather_cat(pc.reshape(-1, 1))
        lm, lc = gm.index_select(0, best), gc.index_select(0, best)          # autograd history only on the owning rank
        a_r = (pm.reshape(-1, 1) - lm.detach()).abs().mean() + (pc.reshape(-1, 1) - lc.detach()).abs().mean()
        b_all = world * ((gm.detach() - lm).abs().mean() + (gc.detach() - lc).abs().mean())
        return a_r + (b_all - b_all.detach())
    best = torch.argmax(scores)
    lm, lc = pm.index_select(0, best.reshape(1)), pc.index_select(0, best.reshape(1))
    return L1MeanFn.apply(pm, lm.expand_as(pm)) + L1MeanFn.apply(pc, lc.expand_as(pc))


class L_TV(torch.nn.Module):
    def __init__(self, TVLoss_weight=1):
        super().__init__()
        self.TVLoss_weight = TVLoss_weight

    def forward(self, x):
        return self.TVLoss_weight * TVFn.apply(x)
