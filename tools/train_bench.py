import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uncltmo_b200 import synth, _lib
from uncltmo_b200.discriminator import SimpleDiscriminator
from uncltmo_b200.generator import UNet
from uncltmo_b200.trainer import GanTrainerStep
from uncltmo_b200.weights import make_discriminator_state_dict, make_generator_state_dict
G_ARGS = (1, 1, "sigmoid", 4, 4, "square_and_square_root", 32, 0, "unet", 0, 0, "none", "none", "relu", 1, "replicate", 2)
PREC = sys.argv[1] if len(sys.argv) > 1 else "fp32"
netG = UNet(*G_ARGS, up_mode=0, precision=PREC).cuda().train(); netG.load_state_dict(make_generator_state_dict())
netD = SimpleDiscriminator(256, 1, 16, "none", "none", 0, 0).cuda().train(); netD.load_state_dict(make_discriminator_state_dict())
optG = torch.optim.Adam([p for p in netG.parameters() if p.requires_grad], lr=1e-5, betas=(0.5, 0.999))
optD = torch.optim.Adam(netD.parameters(), lr=1.5e-5, betas=(0.5, 0.999))
tr = GanTrainerStep(netG, netD, optG, optD)
B = 8
hdr = torch.from_numpy(synth.normalised_batch(2 * B, seed=4)).reshape(B, 2, 1, 256, 256).cuda()
pos = torch.from_numpy(synth.ldr_batch(2 * B, seed=5)).reshape(B, 2, 1, 256, 256).cuda()
neg = torch.from_numpy(synth.ldr_batch(2 * B, seed=6)).reshape(B, 2, 1, 256, 256).cuda()
for _ in range(3): tr.step(hdr, None, pos, neg, 0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time(); e0.record()
K = 5
for _ in range(K): tr.step(hdr, None, pos, neg, 0)
e1.record(); torch.cuda.synchronize()
print(PREC, "train step: %.2f ms device, %.2f ms wall  -> %.2f steps/s" % (e0.elapsed_time(e1) / K, (time.time() - t0) * 1e3 / K, K / (e0.elapsed_time(e1) / 1e3)))
_lib.start_call_timing(); tr.step(hdr, None, pos, neg, 0); rec = _lib.stop_call_timing()
agg = {}
for n, ms in rec: agg[n] = agg.get(n, 0) + ms
for n, ms in sorted(agg.items(), key=lambda kv: -kv[1])[:14]: print("  %-32s %8.3f ms" % (n, ms))
print("  total kernels", sum(agg.values()))
print("  top-level count", len(rec))
# the whole iteration as one CUDA graph (what bench.py times)
try:
    gG = UNet(*G_ARGS, up_mode=0, precision=PREC).cuda().train(); gG.load_state_dict(make_generator_state_dict())
    gD = SimpleDiscriminator(256, 1, 16, "none", "none", 0, 0).cuda().train(); gD.load_state_dict(make_discriminator_state_dict())
    oG = torch.optim.Adam([p for p in gG.parameters() if p.requires_grad], lr=1e-5, betas=(0.5, 0.999), capturable=True, fused=True)
    oD = torch.optim.Adam(gD.parameters(), lr=1.5e-5, betas=(0.5, 0.999), capturable=True, fused=True)
    tg = GanTrainerStep(gG, gD, oG, oD)
    tg.capture(hdr, None, pos, neg, 0)
    for _ in range(3): tg.replay(hdr, None, pos, neg, 0)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20): tg.replay(hdr, None, pos, neg, 0)
    e1.record(); torch.cuda.synchronize()
    print(PREC, "graph replay: %.3f ms/step -> %.1f steps/s" % (e0.elapsed_time(e1) / 20, 20 / (e0.elapsed_time(e1) / 1e3)))
except Exception as e:
    import traceback; traceback.print_exc()
try:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        tr.step(hdr, None, pos, neg, 0)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
except Exception as e:  # CUPTI may be unavailable on the box
    print("profiler unavailable:", e)
