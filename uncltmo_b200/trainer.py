"""The training step of GanTrainerImg (image TMO) and GanTrainer (video TMO), mirrored on the B200 path.

Reference: GanTrainerImg.py:200-217 (train_D), :231-260 (D_real_fake_pass), :262-292 (train_G), :302-339
(update_g_d_loss, the epoch-dependent loss schedule), :452-461 (update_struct_loss); the video trainer
GanTrainer.py:233-262, :264-300, :310-349 is the same step with 5-D clip batches and the recurrent generator (pass
hdr_input as [B,T,1,256,256] and a UNetVideo).  Differences, none of which
changes a gradient that reaches an optimizer:
  * the D-step generator forward and the discriminator passes over real images run without an autograd graph
    (the reference builds the graphs and throws them away; `D(real_neg)` of the D step is computed and never used
    there - GanTrainerImg.py:237 - and is skipped here);
  * errG_d and errG_struct are summed and back-propagated once instead of twice with retain_graph=True;
  * the generator step back-propagates through the discriminator without forming the discriminator's own parameter
    gradients (the reference accumulates them and zeroes them unread at the start of the next train_D);
  * the TMQI naturalness that picks positives / negatives / the pseudo label is computed on the device
    (the reference makes 80 host numpy calls per step), so a step has no host synchronisation;
  * the `epoch > epoch_step2` branch uses L_TV from GanTrainer.py:669-682 (the image trainer references an undefined
    name there).

Multi-GPU (one process per GPU, torch.distributed initialised): every rank takes an equal slice of the global batch.
Per-sample-mean losses are scaled by 1/world and gradients are SUMMED over ranks with bucketed NCCL all-reduces
(uncltmo_b200.dist.GradientBuckets) - the reference's counterpart is nn.DataParallel.  The all-pairs contrastive loss
sees the all-gathered logits of the global batch.  The TMQI selections of infoNCE2 / pseudo_label_loss are made over the
GLOBAL batch (all-gather of the scores; the chosen feature maps / statistics are broadcast from the rank that holds them
and their gradient is returned to it - uncltmo_b200.losses, dist.BroadcastRowsFn), as under nn.DataParallel.
The bf16 image path keeps all generator gradients in one flat buffer and sums it with ONE all-reduce, no flattening copy.
"""
import torch
import torch.distributed as dist

from . import losses
from .dist import GradientBuckets, all_gather_cat
from .struct_loss import StructLoss


# stream priorities of the captured iteration (lower = more urgent; clamped to the device's range): the main chain, the
# generator's forward / data-gradient chain, everything else (weight gradients, loss branches, gradient-free D passes)
PRIO_MAIN, PRIO_G = -3, -2


class GanTrainerStep:
    def __init__(self, netG, netD, optimizerG, optimizerD, loss_g_d_factor=0.1, struct_loss_factor=1.0,
                 adv_weight_list=(0.2, 0.2, 0.2), pyramid_weight_list=(1.0, 1.0, 1.0), epoch_step1=6, epoch_step2=9):
        from .generator import UNetVideo
        self.netG, self.netD = netG, netD
        self.video = isinstance(netG, UNetVideo)
        self.optimizerG, self.optimizerD = optimizerG, optimizerD
        self.loss_g_d_factor = loss_g_d_factor
        self.struct_loss_factor = struct_loss_factor
        self.adv_weight_list = [float(a) for a in adv_weight_list]
        self.pyramid_weight_list = [float(p) for p in pyramid_weight_list]
        self.epoch_step1, self.epoch_step2 = epoch_step1, epoch_step2
        self.struct_loss = StructLoss(self.pyramid_weight_list)
        self.errD = self.errG_d = self.errG_struct = None
        self._side = None
        self._side_g = None
        self._loss_streams = []
        # bf16 path: run train_G's generator forward on a second stream NEXT TO the whole D step (it depends on nothing the
        # D step produces - the D step's own generator pass is a separate, gradient-free forward with its own DropPath
        # masks, as in the reference); at 16 images neither chain fills the GPU
        self.overlap_g_forward = True
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        # generator gradients (19.7 MB) are reduced bucket by bucket WHILE backward is still running (hooks)
        self.buckets_G = GradientBuckets(netG.parameters()).install_hooks() if self.world > 1 else None
        self.buckets_D = GradientBuckets(netD.parameters()) if self.world > 1 else None

    def _generator_is_flat(self):
        """True when the generator ran through train_graph (bf16 image path): every p.grad is a view of one flat buffer."""
        fp = getattr(self.netG, "_flat", None)
        return getattr(self.netG, "precision", None) == "bf16" and fp is not None and fp.grads_attached()

    @staticmethod
    def _flat(t):
        return t.reshape(-1, t.shape[-3], t.shape[-2], t.shape[-1]).float()

    def _generate(self, hdr_input):
        """Image trainer: G on [B,1,256,256].  Video trainer (GanTrainer.py:241-243, 274-276): G on the 5-D clip batch
        [B,T,1,256,256] (frames of a clip are chained through the recurrent hand-over), outputs flattened to B*T."""
        if self.video:
            if hdr_input.dim() != 5:
                raise ValueError("the video trainer expects [B,T,1,256,256] clips")
            if getattr(self.netG, "precision", None) == "bf16" and hasattr(self.netG, "forward_blocked"):
                fake, fea = self.netG.forward_blocked(hdr_input.float())    # the clip as one autograd node, flat gradients
            else:
                fake, fea = self.netG(hdr_input.float())
            return self._flat(fake), self._flat(fea)
        if getattr(self.netG, "precision", None) == "bf16" and hasattr(self.netG, "forward_blocked"):
            # bf16 path: one autograd node for the whole generator, features stay C8-blocked bf16 (losses.infoNCE2 reads them)
            return self.netG.forward_blocked(self._flat(hdr_input))
        return self.netG(self._flat(hdr_input))

    # ------------------------------------------------------------------ D step
    def train_D(self, hdr_input, real_ldr_pos, real_ldr_neg, epoch):
        self.netD.zero_grad(set_to_none=True)
        with torch.no_grad():
            fake, _ = self._generate(hdr_input)
        # D(real_pos) and D(fake) as ONE discriminator pass over the stacked batch (per-sample network: identical values,
        # half the launches and twice the CTAs per launch); the feature statistics are not needed in the D step
        pos = self._flat(real_ldr_pos)
        logits, _ = self.netD(torch.cat([pos, fake.detach()]), want_features=False)
        d_real_pos, d_fake = logits[:pos.shape[0]], logits[pos.shape[0]:]
        w = self.adv_weight_list[0] * (1.0 if epoch <= self.epoch_step1 else 1e-6)
        self.errD = w * losses.contrastive_D_loss(all_gather_cat(d_real_pos), all_gather_cat(d_fake))
        self.errD.backward()
        self.errD = self.errD.detach()   # keep the value, let the graph (and its AccumulateGrad nodes) go
        if self.buckets_D is not None:
            self.buckets_D.allreduce()
        self.optimizerD.step()
        return self.errD

    # ------------------------------------------------------------------ independent loss branches on their own streams
    def _fork(self, idx, fn, *inputs):
        """Run `fn` (a loss term that depends only on `inputs`) on side stream `idx` while the caller continues.  Only inside a
        CUDA-graph capture: there the terms become parallel branches of the graph (forward AND backward - autograd runs a
        node's backward on its forward's stream), which matters because every term is a chain of 3-30 us kernels; eager
        launches are CPU-bound and would gain nothing.  Returns a handle for `_get`."""
        # (single GPU only: with several ranks some terms contain collectives, which stay on the main stream in program order)
        if self.world > 1 or not torch.cuda.is_current_stream_capturing():
            return (None, fn())
        cur = torch.cuda.current_stream()
        while len(self._loss_streams) <= idx:
            self._loss_streams.append(torch.cuda.Stream())
        s = self._loss_streams[idx]
        s.wait_stream(cur)
        for t in inputs:
            if t is not None:
                t.record_stream(s)
        with torch.cuda.stream(s):
            out = fn()
        return (s, out)

    @staticmethod
    def _get(handle):
        s, out = handle
        if s is not None:
            cur = torch.cuda.current_stream()
            cur.wait_stream(s)
            for t in (out if isinstance(out, (tuple, list)) else (out,)):
                t.record_stream(cur)
        return out

    # ------------------------------------------------------------------ G step
    def image_terms(self, fea_fake, fake, hdr_input, ldr_pos, epoch):
        """The terms of update_g_d_loss that read only the generator's outputs (not the discriminator's): started right
        after the generator forward, each on its own branch (`_fork`), next to the D(fake) pass."""
        if epoch <= self.epoch_step2:
            return {"nce2": self._fork(0, lambda: losses.infoNCE2(fea_fake, fake, hdr_input, "InfoNCE", 1, 1e-2), fea_fake, fake),
                    "l1": self._fork(1, lambda: losses.l1_mean_terms(fake, ldr_pos), fake, ldr_pos),
                    "pseudo": self._fork(2, lambda: losses.pseudo_label_loss(fake, hdr_input), fake)}
        return {"l1": self._fork(1, lambda: losses.l1_mean_terms(fake, ldr_pos), fake, ldr_pos),
                "pseudo": self._fork(2, lambda: losses.pseudo_label_loss(fake, hdr_input), fake),
                "tv": self._fork(0, lambda: losses.L_TV()(fake), fake)}

    def g_d_loss(self, d_fake_bp, d_real_pos_bp, d_fea_fake, d_fea_real_pos, d_fea_real_neg, d_fea_input, fea_fake, fake,
                 hdr_input, ldr_pos, epoch, terms=None):
        """update_g_d_loss (GanTrainerImg.py:302-339) without the backward call."""
        f = self.loss_g_d_factor
        s = 1.0 / self.world   # per-sample-mean terms: local mean * (local / global batch)
        if terms is None:
            terms = self.image_terms(fea_fake, fake, hdr_input, ldr_pos, epoch)
        gathered = losses.contrastive_D_loss(all_gather_cat(d_fake_bp), all_gather_cat(d_real_pos_bp))
        if epoch <= self.epoch_step2:
            first = epoch <= self.epoch_step1
            err = f * (1.0 if first else 1e-6) * gathered
            f = f * s
            err = err + f * 0.5 * losses.infoNCE(d_fea_fake, d_fea_real_pos, d_fea_input, fake, hdr_input, "InfoNCE", 1, 1e-2)
            err = err + f * 0.5 * 0.2 * losses.infoNCE(d_fea_fake, d_fea_real_pos, d_fea_real_neg, fake, hdr_input, "InfoNCE", 1e3, 2)
            err = err + f * (1e-6 if first else 0.5) * self._get(terms["nce2"])
            l_mean, l_con = self._get(terms["l1"])
            err = err + f * (1e-6 if first else 0.5 * 1e2) * l_mean
            err = err + f * (1e-6 if first else 0.5 * 2) * l_con
            err = err + f * 1e-6 * self._get(terms["pseudo"])
        else:
            err = f * 1e-6 * gathered
            f = f * s
            l_mean, _ = self._get(terms["l1"])
            err = err + f * 0.5 * 1e2 * l_mean
            err = err + f * 0.5 * 1e2 * self._get(terms["pseudo"])
            err = err + f * 0.2 * 1e5 * self._get(terms["tv"])
        return err

    def train_G(self, hdr_input, hdr_original_gray_norm, real_ldr_pos, real_ldr_neg, epoch, generated=None, early=None):
        """generated: (fake, fea_fake) of THIS step's generator forward when `step` already ran it next to the D step;
        early: the (terms, struct) branch handles `step` forked from that forward."""
        self.netG.zero_grad(set_to_none=True)
        hdr = self._flat(hdr_input)
        pos, neg = self._flat(real_ldr_pos), self._flat(real_ldr_neg)
        # the three discriminator passes over real / input images carry no gradient and do not depend on the generator.
        # Inside a CUDA-graph capture they are forked onto a side stream next to the generator forward (small,
        # latency-bound kernels fill idle SMs: -0.5 ms per replayed step); eager launches are CPU-bound, so there the
        # extra stream only costs time and the passes stay in line.
        cur = torch.cuda.current_stream()
        fork = torch.cuda.is_current_stream_capturing()
        if fork:
            if self._side is None:
                self._side = torch.cuda.Stream()
            self._side.wait_stream(cur)
        with torch.cuda.stream(self._side if fork else cur), torch.no_grad():
            nb = pos.shape[0]
            lg, fe = self.netD(torch.cat([pos, neg, hdr]))      # one pass over the three gradient-free batches
            d_real_pos_bp, d_fea_real_pos, d_fea_real_neg, d_fea_input = lg[:nb], fe[:nb], fe[nb:2 * nb], fe[2 * nb:]
        fake, fea_fake = generated if generated is not None else self._generate(hdr_input)
        if fork:
            cur.wait_stream(self._side)
            for t in (d_real_pos_bp, d_fea_real_pos, d_fea_real_neg, d_fea_input):
                t.record_stream(cur)
        # the loss terms that only read the generator's outputs start now, as parallel branches next to D(fake)
        if early is not None:
            terms, struct = early
        else:
            terms = self.image_terms(fea_fake, fake, hdr, pos, epoch)
            struct = self._fork(3, lambda: self.struct_loss(fake, None, hdr, self.pyramid_weight_list), fake, hdr) \
                if self.struct_loss_factor else None
        # back-propagation THROUGH the discriminator: its own parameter gradients would be zeroed unread by the next train_D
        # (GanTrainerImg.py:201), so they are not computed (a third of the discriminator backward)
        d_params = [p for p in self.netD.parameters() if p.requires_grad]
        for p in d_params:
            p.requires_grad_(False)
        try:
            d_fake_bp, d_fea_fake = self.netD(fake)
        finally:
            for p in d_params:
                p.requires_grad_(True)
        self.errG_d = self.g_d_loss(d_fake_bp, d_real_pos_bp, d_fea_fake, d_fea_real_pos, d_fea_real_neg, d_fea_input,
                                    fea_fake, fake, hdr, pos, epoch, terms)
        total = self.errG_d
        if struct is not None:
            self.errG_struct = (self.struct_loss_factor / self.world) * self._get(struct)
            total = total + self.errG_struct
        flat_mode = self.world > 1 and self._generator_is_flat()
        if self.buckets_G is not None and not flat_mode:
            self.buckets_G.arm()
        total.backward()
        del total
        self.errG_d = self.errG_d.detach()
        if self.errG_struct is not None:
            self.errG_struct = self.errG_struct.detach()
        if flat_mode:
            dist.all_reduce(self.netG._flat.grad, op=dist.ReduceOp.SUM)    # the persistent flat gradient buffer, in place
        elif self.buckets_G is not None:
            self.buckets_G.finish()
        self.optimizerG.step()
        return self.errG_d, self.errG_struct

    def step(self, hdr_input, hdr_original_gray_norm, real_ldr_pos, real_ldr_neg, epoch):
        """One iteration of GanTrainer.train_epoch's loop body (GanTrainerImg.py:178-186)."""
        generated = None
        if self.overlap_g_forward and getattr(self.netG, "precision", None) == "bf16" and hasattr(self.netG, "forward_blocked") \
                and torch.is_grad_enabled():
            from .train_graph import flat_params
            cur = torch.cuda.current_stream()
            flat_params(self.netG).pack()          # both generator passes read the packed weights: pack before the fork
            if self._side_g is None:
                self._side_g = torch.cuda.Stream(priority=PRIO_G)
            self._side_g.wait_stream(cur)
            early = None
            with torch.cuda.stream(self._side_g):
                generated = self._generate(hdr_input)
                if self.world == 1 and torch.cuda.is_current_stream_capturing():
                    # the loss terms that read only the generator's outputs branch off here, next to the D step as well
                    fake, fea_fake = generated
                    hdr, pos = self._flat(hdr_input), self._flat(real_ldr_pos)
                    terms = self.image_terms(fea_fake, fake, hdr, pos, epoch)
                    struct = self._fork(3, lambda: self.struct_loss(fake, None, hdr, self.pyramid_weight_list), fake, hdr) \
                        if self.struct_loss_factor else None
                    early = (terms, struct)
        self.train_D(hdr_input, real_ldr_pos, real_ldr_neg, epoch)
        if generated is not None:
            cur.wait_stream(self._side_g)
            for t in generated:
                if t is not None:
                    t.record_stream(cur)
            return self.train_G(hdr_input, hdr_original_gray_norm, real_ldr_pos, real_ldr_neg, epoch, generated, early)
        return self.train_G(hdr_input, hdr_original_gray_norm, real_ldr_pos, real_ldr_neg, epoch)

    # ------------------------------------------------------------------ CUDA-graph replay of the whole iteration
    def _invalidate_packed(self):
        self.netG._packed = None
        flat = getattr(self.netG, "_flat", None)
        if flat is not None:
            flat._packed_version = None

    def _branch(self, epoch):
        return 0 if epoch <= self.epoch_step1 else (1 if epoch <= self.epoch_step2 else 2)

    def capture(self, hdr_input, hdr_original_gray_norm, real_ldr_pos, real_ldr_neg, epoch, warmup=3):
        """Capture one whole iteration - train_D, train_G, both optimizer steps and (multi-GPU) the gradient
        all-reduces - into a CUDA graph for the loss schedule `epoch` falls in.  The step makes ~1000 kernel launches of
        10 us each and has no host synchronisation, so a replay removes the launch-bound floor (it matters most when the
        batch is split over several GPUs).  `warmup` REAL iterations run on the given batch first (lazy optimizer state
        and lazily initialised kernels must exist before the capture).  Optimizers must be built with capturable=True.
        Input shapes are frozen; `replay` copies new batches into the captured buffers."""
        for opt in (self.optimizerG, self.optimizerD):
            if not all(g.get("capturable", False) for g in opt.param_groups):
                raise ValueError("GanTrainerStep.capture needs optimizers constructed with capturable=True")
        if not hasattr(self, "_graphs"):
            self._graphs = {}
        static = [hdr_input.clone(), real_ldr_pos.clone(), real_ldr_neg.clone()]
        self.errD = self.errG_d = self.errG_struct = None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self.step(static[0], None, static[1], static[2], epoch)
        torch.cuda.current_stream().wait_stream(side)
        self.netG.zero_grad(set_to_none=True)
        self.netD.zero_grad(set_to_none=True)
        self._invalidate_packed()     # the weight re-layout of the D-step generator pass must be part of the graph
        graph = torch.cuda.CUDAGraph()
        # the capture stream carries the iteration's critical chain (D step -> D(fake) -> backward): highest priority, so that
        # its kernels are scheduled ahead of the side branches' (graph kernel nodes inherit the capturing stream's priority)
        if getattr(self, "_capture_stream", None) is None:
            self._capture_stream = torch.cuda.Stream(priority=PRIO_MAIN)
        with torch.cuda.graph(graph, stream=self._capture_stream):
            err_g, err_s = self.step(static[0], None, static[1], static[2], epoch)
        self._invalidate_packed()
        self._graphs[self._branch(epoch)] = (graph, static, (self.errD, err_g, err_s))
        return graph

    def replay(self, hdr_input, hdr_original_gray_norm, real_ldr_pos, real_ldr_neg, epoch):
        """`step` through the captured graph: returns (errG_d, errG_struct); `self.errD` as after `step`."""
        graph, static, (err_d, err_g, err_s) = self._graphs[self._branch(epoch)]
        for dst, src in zip(static, (hdr_input, real_ldr_pos, real_ldr_neg)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        graph.replay()
        self._invalidate_packed()     # parameters changed behind the version counters the packing caches key on
        self.errD, self.errG_d, self.errG_struct = err_d, err_g, err_s
        return err_g, err_s
