"""Loss methods of GanTrainer / GanTrainerImg with the reference call signatures, forward on sm_100a kernels.

Reference: GanTrainerImg.py:219-229 (contrastive_D_loss), :410-439 (nce), :370-382 (infoNCE), :308-313 (mean /
contrast L1 terms), GanTrainer.py:669-682 (L_TV).  `infoNCE2` / `pseudo_label_loss` choose their positives with a
host TMQI score (GanTrainerImg.py:341-408); here the selection is passed in as indices (`nce_from_indices`) - the
on-device TMQI-naturalness score is the next row of SURVEY.md §8(f2).
"""
import torch

from ._lib import call
from .features import plane_mean_contrast


def _scalar(dev):
    return torch.empty(1, device=dev, dtype=torch.float32)


def contrastive_D_loss(real_logits, fake_logits):
    r = real_logits.reshape(-1).contiguous().float()
    f = fake_logits.reshape(-1).contiguous().float()
    out = _scalar(r.device)
    call("uncl_contrastive_d_loss", r, f, r.numel(), out)
    return out[0]


def nce(fea_anchor, feas_positive, feas_negative, cl_loss_type, k, constant):
    if cl_loss_type != "InfoNCE" or len(feas_positive) != 1 or len(feas_negative) != 1:
        raise NotImplementedError("only the shipped call pattern (InfoNCE, one positive, one negative) is built")
    a = fea_anchor.contiguous().float()
    b, c, h, w = a.shape

    def prep(t):
        if t.shape[0] == 1 or (t.stride(0) == 0):
            return t[:1].contiguous().float(), 0
        return t.contiguous().float(), c * h * w

    p, ps = prep(feas_positive[0])
    n, ns = prep(feas_negative[0])
    logits = torch.empty(2 * b, device=a.device, dtype=torch.float32)
    out = _scalar(a.device)
    call("uncl_nce_fwd", a, p, ps, n, ns, b, c, h * w, float(k), float(constant), logits, out)
    return out[0]


def infoNCE(fea_fake, fea_real, fea_neg, fake, hdr_input, cl_loss_type, k, constant):
    return nce(fea_fake, [fea_real], [fea_neg], cl_loss_type, k, constant)


def nce_from_indices(fea_fake, pos_index, neg_index, cl_loss_type, k, constant):
    """infoNCE2 after the selection: positive / negative = one sample of the batch, broadcast (GanTrainerImg.py:400-405)."""
    return nce(fea_fake, [fea_fake[pos_index:pos_index + 1]], [fea_fake[neg_index:neg_index + 1]], cl_loss_type, k, constant)


def l1_mean_terms(fake, ldr):
    """(L1 of per-image means, L1 of per-image mean local variance): GanTrainerImg.py:308-313."""
    fm, fc = plane_mean_contrast(fake)
    lm, lc = plane_mean_contrast(ldr)
    o1, o2 = _scalar(fake.device), _scalar(fake.device)
    call("uncl_l1_mean", fm.reshape(-1), lm.reshape(-1), fm.numel(), o1)
    call("uncl_l1_mean", fc.reshape(-1), lc.reshape(-1), fc.numel(), o2)
    return o1[0], o2[0]


class L_TV(torch.nn.Module):
    def __init__(self, TVLoss_weight=1):
        super().__init__()
        self.TVLoss_weight = TVLoss_weight

    def forward(self, x):
        x = x.contiguous().float()
        b, c, h, w = x.shape
        scratch = torch.empty(2, device=x.device, dtype=torch.float32)
        out = _scalar(x.device)
        call("uncl_tv_loss", x, b, c, h, w, scratch, out)
        return self.TVLoss_weight * out[0]
