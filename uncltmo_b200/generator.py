"""Drop-in generator modules for the UnCLTMO `unet_multi_filters` U-Net (image and video variants).

Same constructor / forward signatures and `state_dict` keys as the reference classes
(models/unet_multi_filters/Unet_singleFrame.py:101-213, Unet.py:135-289), but
every operator runs as an sm_100a kernel from libuncltmo_b200.so.  Only the shipped hyper-parameters are
built (SURVEY.md §5 "config"); anything else raises at construction - there is no fallback.

precision:
  "fp32"    - CUDA-core fp32 kernels (generator rel-L2 <= 1e-4 vs the reference).
  "bf16"    - bf16 activations; the 3x3 convolutions run on tcgen05 tensor cores (conv_tc.cu), fp32 accumulation;
              the KNN graph of the bottleneck stays fp32 (rel-L2 <= 1e-2).
  "fp32_tc" - the exact path ON the tensor cores (inference): fp32 activations, every 3x3 / k2 s2 convolution as a
              three-term bf16 split x_hi.w_hi + x_hi.w_lo + x_lo.w_hi with fp32 accumulation (~2^-16 per product instead
              of bf16's 2^-9; rel-L2 <= 1e-4), through the same tcgen05 kernels with 3x the input channels; the graph
              block stays on the fp32 CUDA-core kernels.
"""
import torch
import torch.nn as nn

from . import _lib, packing
from ._lib import ACT_GELU, ACT_NONE, ACT_RELU, call

SHIPPED = dict(n_channels=1, output_dim=1, last_layer="sigmoid", depth=4, layer_factor=4,
               con_operator="square_and_square_root", filters=32, network="unet", unet_norm="none",
               stretch_g="none", activation="relu", padding_mode="replicate", convtranspose_kernel=2)


class _WB(nn.Module):
    """weight (+bias) holder that contributes `<name>.weight` / `<name>.bias` to the state_dict."""

    def __init__(self, wshape, bshape=None):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(wshape))
        if bshape is not None:
            self.bias = nn.Parameter(torch.zeros(bshape))


class _Holder(nn.Module):
    pass


def _double(ci, co, second_transposed=False):
    h = _Holder()
    h.conv = _WB((co, ci, 3, 3), (co,))
    h.conv1 = _WB((co, co, 3, 3), (co,))
    h.second_transposed = second_transposed
    return h


class _GeneratorBase(nn.Module):
    def __init__(self, n_channels, output_dim, last_layer, depth, layer_factor, con_operator, filters, bilinear,
                 network, dilation, to_crop, unet_norm, stretch_g, activation, doubleConvTranspose, padding_mode,
                 convtranspose_kernel, up_mode=True, recurrent_ch_ratio=1 / 32, precision="fp32"):
        super().__init__()
        got = dict(n_channels=n_channels, output_dim=output_dim, last_layer=last_layer, depth=depth,
                   layer_factor=layer_factor, con_operator=con_operator, filters=filters, network=network,
                   unet_norm=unet_norm, stretch_g=stretch_g, activation=activation, padding_mode=padding_mode,
                   convtranspose_kernel=convtranspose_kernel)
        bad = {k: v for k, v in got.items() if SHIPPED[k] != v}
        if bad or bilinear or up_mode or not doubleConvTranspose:
            raise NotImplementedError("uncltmo_b200 builds the shipped generator configuration only; unsupported: %r "
                                      "(bilinear=%r up_mode=%r doubleConvTranspose=%r)"
                                      % (bad, bilinear, up_mode, doubleConvTranspose))
        if precision not in ("fp32", "bf16", "fp32_tc"):
            raise ValueError("precision must be 'fp32', 'bf16' or 'fp32_tc'")
        self.precision = precision
        # bf16 inference: skip^2 / sqrt(skip + eps) of the two largest levels are built inside the consuming conv instead of
        # being written by the producer (uncl_conv3x3_tc_skipcat): 1.1 GB less DRAM traffic per 1080p frame.  Kernel for kernel
        # the time is neutral (the in-kernel transform shares the shared-memory port the N' = 96 MMAs saturate: producers -76
        # us, consumers +70 us), but back-to-back frames run at the 1 kW power cap, where the saved traffic buys clock:
        # 485 -> 496 frames/s sustained (tools/fused_skip_sustained.py).  False = materialised planes.
        self.fused_skip = True
        # bf16 inference: the narrow layers at 122..256 pixels (inc.conv1, down0.conv*, up2.conv*, up3.conv*) run in the row kernel
        # (conv_tc_rows.cu: ky taps merged into N, every input row through shared memory once).  False = the older kernels.
        self.row_kernel = True
        # ... and inc.conv (1 -> 32) is computed inside inc.conv1's launch: its 32-channel output never reaches memory
        # (uncl_conv_first_conv3x3_tc_rows).  False = uncl_conv_first + the row kernel.
        self.fused_first = True
        self.to_crop = to_crop
        self.depth = depth
        self.recurrent_ch_ratio = recurrent_ch_ratio
        f = filters
        self.inc = _Holder()
        self.inc.conv = _double(1, f)
        self.inc.conv.conv = _WB((f, 1, 3, 3), (f,))
        self.down_path = nn.ModuleList()
        ch = f
        for i in range(3):
            d = _Holder()
            d.mpconv = nn.ModuleList([nn.Identity(), _double(ch, 2 * ch)])
            self.down_path.append(d)
            ch *= 2
        d = _Holder()
        d.mpconv = nn.ModuleList([nn.Identity(), _double(ch, ch, second_transposed=True)])
        self.down_path.append(d)
        # bottleneck graph block
        self.gcn = _Holder()
        self.gcn.pos_embed = nn.Parameter(torch.zeros(1, ch, 12, 12))
        grapher = _Holder()
        grapher.relative_pos = nn.Parameter(torch.zeros(1, 144, 144), requires_grad=False)
        grapher.fc1 = nn.ModuleList([_WB((ch, ch, 1, 1), (ch,))])
        grapher.graph_conv = _Holder()
        grapher.graph_conv.gconv = _Holder()
        grapher.graph_conv.gconv.nn = nn.ModuleList([_WB((2 * ch, 2 * ch // 4, 1, 1), (2 * ch,))])
        grapher.fc2 = nn.ModuleList([_WB((ch, 2 * ch, 1, 1), (ch,))])
        ffn = _Holder()
        ffn.fc1 = nn.ModuleList([_WB((ch, ch, 1, 1), (ch,))])
        ffn.fc2 = nn.ModuleList([_WB((ch, ch, 1, 1), (ch,))])
        self.gcn.module = nn.ModuleList([nn.ModuleList([grapher, ffn])])
        self.drop_path_prob = 0.05  # GCNBlock: dpr = linspace(0.05, 0.1, 1)[0]  (Unet_singleFrame.py:63)
        self.up_path = nn.ModuleList()
        for i in range(4):
            co = f if i >= 2 else ch // 2
            u = _Holder()
            u.up = _WB((ch, ch, 2, 2), (ch,))
            u.conv = _Holder()
            u.conv.conv = _WB((4 * ch, co, 3, 3), (co,))
            u.conv.conv1 = _WB((co, co, 3, 3), (co,))
            self.up_path.append(u)
            ch //= 2
        self.outc = _Holder()
        self.outc.conv = _WB((1, f, 1, 1), (1,))
        self._packed = None
        self._packed_key = None
        from .weights import relative_pos_table
        with torch.no_grad():
            grapher.relative_pos.copy_(relative_pos_table(8 * f, 12))

    # ------------------------------------------------------------------ packing
    def _pack_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters()) + (self.precision,)

    def packed(self):
        flat = getattr(self, "_flat", None)
        if flat is not None and self.precision == "bf16" and flat.is_current():
            # training keeps the parameters in one flat buffer: every operand comes from ONE gather launch per update
            return flat.pack().P()
        key = self._pack_key()
        if self._packed is None or key != self._packed_key:
            self._packed = self._pack()
            self._packed_key = key
        return self._packed

    def _pack(self):
        tc = self.precision == "bf16"
        split = self.precision == "fp32_tc"
        P = {}

        def conv(name, m, transposed):
            w9 = packing.conv3x3_taps(m.weight.detach(), transposed)
            wp = packing.conv3x3_tc(w9) if tc else (packing.conv3x3_tc_split(w9) if split else w9)
            P[name] = (wp, m.bias.detach().float().contiguous())
            if tc and name in self._ROW_LAYERS:
                ci, co, derive = self._ROW_LAYERS[name]
                if packing.conv3x3_tc_rows_plan(1, ci, 64, 64, 0, derive, co=co)[0] == 1:
                    P[name + "_rows"] = packing.conv3x3_tc_rows(w9)

        with torch.no_grad():
            P["inc0"] = (packing.conv_first(self.inc.conv.conv.weight.detach()), self.inc.conv.conv.bias.detach().float().contiguous())
            if tc:
                P["inc0_rows"] = packing.conv_first_rows(self.inc.conv.conv.weight.detach(), self.inc.conv.conv.bias.detach())
            conv("inc1", self.inc.conv.conv1, False)
            for i in range(4):
                blk = self.down_path[i].mpconv[1]
                conv("d%d_0" % i, blk.conv, False)
                conv("d%d_1" % i, blk.conv1, i == 3)
            g, ffn = self.gcn.module[0][0], self.gcn.module[0][1]
            P["pos"] = packing.blocked_param(self.gcn.pos_embed.detach())
            P["relpos"] = g.relative_pos.detach().reshape(144, 144).float().contiguous()
            for name, m, groups in (("g_fc1", g.fc1[0], 1), ("g_gconv", g.graph_conv.gconv.nn[0], 4),
                                    ("g_fc2", g.fc2[0], 1), ("f_fc1", ffn.fc1[0], 1), ("f_fc2", ffn.fc2[0], 1)):
                pack = packing.pointwise_tc if tc and name != "g_fc1" else packing.pointwise
                P[name] = (pack(m.weight.detach(), groups), m.bias.detach().float().contiguous())
            if tc:   # fc1 on the tensor cores as a three-term bf16 split GEMM (it feeds the KNN selection: ~2^-16, not 2^-9)
                P["g_fc1_split"] = (packing.pointwise_tc_split(g.fc1[0].weight.detach()), P["g_fc1"][1])
            for i in range(4):
                u = self.up_path[i]
                wu = u.up.weight.detach()
                P["u%d_up" % i] = (packing.convT2x2_tc(wu) if tc else (packing.convT2x2_tc_split(wu) if split else packing.convT2x2(wu)),
                                   u.up.bias.detach().float().contiguous())
                conv("u%d_0" % i, u.conv.conv, True)
                conv("u%d_1" % i, u.conv.conv1, True)
            P["outc"] = (self.outc.conv.weight.detach().reshape(-1).float().contiguous(),
                         self.outc.conv.bias.detach().float().contiguous())
        return P

    # layers the row kernel takes in bf16 inference: name -> (logical C_in, C_out, fused skip operators)
    _ROW_LAYERS = {"inc1": (32, 32, False), "d0_0": (32, 64, False), "d0_1": (64, 64, False), "u2_0": (256, 32, True),
                   "u2_1": (32, 32, False), "u3_0": (128, 32, True), "u3_1": (32, 32, False)}

    # ------------------------------------------------------------------ one frame through the network
    def _conv3(self, P, name, src, src_stride, dst, dst_stride, n, ci, h, w, co, pad, emit_skip=0, fuse=None):
        wt, b = P[name]
        rows = P.get(name + "_rows") if (self.precision == "bf16" and self.row_kernel) else None
        if rows is not None:
            ow, ob, out_img, out_logit = fuse if fuse is not None else (None, None, None, None)
            call("uncl_conv3x3_tc_rows", src, src_stride, rows, wt, b, dst, dst_stride, n, ci, h, w, co, pad, ACT_RELU, emit_skip,
                 0 if fuse is None else 1, ow, ob, out_img, out_logit)
        elif self.precision == "bf16":
            if fuse is None:
                call("uncl_conv3x3_tc", src, src_stride, wt, b, dst, dst_stride, _lib.BF16, n, ci, h, w, co, pad, ACT_RELU,
                     emit_skip, 0, None, None, None, None)
            else:
                ow, ob, out_img, out_logit = fuse
                call("uncl_conv3x3_tc", src, src_stride, wt, b, dst, dst_stride, _lib.BF16, n, ci, h, w, co, pad, ACT_RELU,
                     emit_skip, 1, ow, ob, out_img, out_logit)
        elif self.precision == "fp32_tc":
            xs = torch.empty((n, 3 * ci // 8, h, w, 8), device=src.device, dtype=torch.bfloat16)
            call("uncl_split_bf16", src, src_stride, xs, xs.stride(0), n, ci, h * w)
            call("uncl_conv3x3_tc", xs, xs.stride(0), wt, b, dst, dst_stride, _lib.F32, n, 3 * ci, h, w, co, pad, ACT_RELU,
                 emit_skip, 0, None, None, None, None)
        else:
            call("uncl_conv3x3_simt", src, src_stride, wt, b, dst, dst_stride, n, ci, h, w, co, pad, ACT_RELU,
                 emit_skip, _lib.F32)

    def _run_frame(self, x, prev=None, droppath_scale=None, want_features=True, want_logit=False, keep=None):
        """x: [N,1,H,W] fp32 CUDA contiguous with H=W=256.  prev: recurrent state of the previous frame (video).
        Returns (out [N,1,256,256] fp32, up_x blocked tensor or None, logit or None, state list)."""
        if x.dim() != 4 or x.shape[1] != 1 or x.shape[2] != 256 or x.shape[3] != 256:
            raise ValueError("the generator only accepts [N,1,256,256] inputs (pos_embed is a fixed 12x12 grid, "
                             "Unet_singleFrame.py:66,94); got %s" % (tuple(x.shape),))
        if not x.is_cuda:
            raise RuntimeError("uncltmo_b200 has no CPU path: move the input to a CUDA device")
        P = self.packed()
        x = x.contiguous().float()
        n = x.shape[0]
        dev = x.device
        tdt = torch.bfloat16 if self.precision == "bf16" else torch.float32
        dt = _lib.DTYPE_OF[tdt]

        def buf(c, h, w):
            return torch.empty((n, c // 8, h, w, 8), device=dev, dtype=tdt)

        def st(t):
            return t.stride(0)

        f = 32
        # encoder.  cat[i] is the concat buffer of up stage 3-i: [skip | upsampled | skip^2 | sqrt(skip)].  On the bf16
        # inference path the two largest levels (252^2, 124^2: 70 % of the concat bytes) keep only [skip | upsampled]: the
        # consuming conv builds skip^2 / sqrt(skip + eps) in shared memory (uncl_conv3x3_tc_skipcat, unet_parts.py:311-332)
        sizes = [(f, 252), (2 * f, 122), (4 * f, 57), (8 * f, 24)]
        nfused = 2 if (self.precision == "bf16" and keep is None and self.fused_skip) else 0
        cat = [buf((2 if i < nfused else 4) * c, s, s) for i, (c, s) in enumerate(sizes)]
        if (self.precision == "bf16" and self.row_kernel and self.fused_first and keep is None and "inc1_rows" in P
                and "inc0_rows" in P):
            call("uncl_conv_first_conv3x3_tc_rows", x, x.stride(0), P["inc0_rows"], P["inc1_rows"], P["inc1"][1],
                 cat[0], st(cat[0]), n, 256, 256, ACT_RELU, 0 if nfused > 0 else 1)
        else:
            a0 = buf(f, 254, 254)
            call("uncl_conv_first", x, P["inc0"][0], P["inc0"][1], a0, st(a0), n, 256, 256, f, ACT_RELU, dt)
            self._conv3(P, "inc1", a0, st(a0), cat[0], st(cat[0]), n, f, 254, 254, f, 0, emit_skip=0 if nfused > 0 else 1)
        state = [cat[0]]  # tensors whose first C/32 channels feed the next frame (Unet.py:229,251,264,272)
        cur, cur_c, cur_s = cat[0], f, 252
        for i in range(4):
            ps = cur_s // 2
            pooled = buf(cur_c, ps, ps)
            pv = prev[i] if prev is not None else None
            call("uncl_maxpool2", cur, st(cur), pv, st(pv) if pv is not None else 0, cur_c // 32 if pv is not None else 0,
                 pooled, st(pooled), n, cur_c, cur_s, cur_s, dt)
            co = cur_c * 2 if i < 3 else cur_c
            mid = buf(co, ps - 2, ps - 2)
            self._conv3(P, "d%d_0" % i, pooled, st(pooled), mid, st(mid), n, cur_c, ps, ps, co, 0)
            if i < 3:
                dst = cat[i + 1]
                self._conv3(P, "d%d_1" % i, mid, st(mid), dst, st(dst), n, co, ps - 2, ps - 2, co, 0,
                            emit_skip=0 if i + 1 < nfused else 1)
                cur, cur_c, cur_s = dst, co, ps - 4
            else:
                x4 = buf(co, ps, ps)
                self._conv3(P, "d3_1", mid, st(mid), x4, st(x4), n, co, ps - 2, ps - 2, co, 2)
                cur, cur_c, cur_s = x4, co, ps
            state.append(cur)
        # bottleneck graph block (fp32 internals)
        C = cur_c

        def f32buf(c):
            return torch.empty((n, c // 8, 144, 8), device=dev, dtype=torch.float32)

        s0 = droppath_scale[0] if droppath_scale is not None else None
        s1 = droppath_scale[1] if droppath_scale is not None else None
        idx = torch.empty((n, 144, 9), device=dev, dtype=torch.int32) if keep is not None else None
        x0, y = f32buf(C), f32buf(C)
        gout = buf(C, 12, 12)
        # fc1 feeds the KNN selection (a discrete choice): fp32 on the CUDA cores in the exact path; in the bf16 path a
        # three-term bf16 split GEMM on the tensor cores (x_hi.w_hi + x_hi.w_lo + x_lo.w_hi, fp32 accumulation and output:
        # ~2^-16 relative, far below the 2^-9 of the bf16 activations it reads)
        if self.precision == "bf16" and "g_fc1_split" in P:
            xs = torch.empty((n, 3 * C // 8, 144, 8), device=dev, dtype=torch.bfloat16)
            call("uncl_gcn_add_pos_split", cur, st(cur), P["pos"], x0, xs, n, C, dt)
            call("uncl_pw_conv_tc", xs, st(xs), P["g_fc1_split"][0], P["g_fc1_split"][1], None, 0, _lib.F32, None, y, st(y),
                 _lib.F32, n, 3 * C, C, 1, 12, 12, ACT_NONE)
        else:
            call("uncl_gcn_add_pos", cur, st(cur), P["pos"], x0, n, C, dt)
            call("uncl_pw_conv", x0, P["g_fc1"][0], P["g_fc1"][1], None, None, y, st(y), n, C, C, 1, 144, ACT_NONE, _lib.F32)
        if self.precision == "bf16":
            # the other four 1x1 convs (92 % of the block's MACs) run as tensor-core GEMMs on bf16 tensors
            def b16buf(c):
                return torch.empty((n, c // 8, 144, 8), device=dev, dtype=torch.bfloat16)

            z, z2, x1, f1 = b16buf(2 * C), b16buf(2 * C), b16buf(C), b16buf(C)
            BF, F32 = _lib.BF16, _lib.F32
            call("uncl_gcn_knn_aggregate", y, P["relpos"], z, BF, idx, n, C)
            call("uncl_pw_conv_tc", z, st(z), P["g_gconv"][0], P["g_gconv"][1], None, 0, BF, None, z2, st(z2), BF, n,
                 2 * C, 2 * C, 4, 12, 12, ACT_GELU)
            call("uncl_pw_conv_tc", z2, st(z2), P["g_fc2"][0], P["g_fc2"][1], x0, st(x0), F32, s0, x1, st(x1), BF, n,
                 2 * C, C, 1, 12, 12, ACT_NONE)
            call("uncl_pw_conv_tc", x1, st(x1), P["f_fc1"][0], P["f_fc1"][1], None, 0, BF, None, f1, st(f1), BF, n,
                 C, C, 1, 12, 12, ACT_GELU)
            call("uncl_pw_conv_tc", f1, st(f1), P["f_fc2"][0], P["f_fc2"][1], x1, st(x1), BF, s1, gout, st(gout), BF, n,
                 C, C, 1, 12, 12, ACT_NONE)
        else:
            z, z2, x1, f1 = f32buf(2 * C), f32buf(2 * C), f32buf(C), f32buf(C)
            call("uncl_gcn_knn_aggregate", y, P["relpos"], z, _lib.F32, idx, n, C)
            call("uncl_pw_conv", z, P["g_gconv"][0], P["g_gconv"][1], None, None, z2, st(z2), n, 2 * C, 2 * C, 4, 144, ACT_GELU, _lib.F32)
            call("uncl_pw_conv", z2, P["g_fc2"][0], P["g_fc2"][1], x0, s0, x1, st(x1), n, 2 * C, C, 1, 144, ACT_NONE, _lib.F32)
            call("uncl_pw_conv", x1, P["f_fc1"][0], P["f_fc1"][1], None, None, f1, st(f1), n, C, C, 1, 144, ACT_GELU, _lib.F32)
            call("uncl_pw_conv", f1, P["f_fc2"][0], P["f_fc2"][1], x1, s1, gout, st(gout), n, C, C, 1, 144, ACT_NONE, dt)
        state.append(gout)
        if keep is not None:
            keep.update(x0=x0, y=y, idx=idx, z=z, z2=z2, x1=x1, f1=f1, gcn=gout, skips=cat, x4=cur)
        # decoder
        up, up_c, up_s = gout, C, 12
        out = torch.empty((n, 1, 256, 256), device=dev, dtype=torch.float32)
        logit = torch.empty_like(out) if want_logit else None
        fused = False
        for i in range(4):
            cb = cat[3 - i]
            sk_c, sk_s = sizes[3 - i]
            # ConvTranspose k2 s2 into the second channel group of the concat buffer
            dst = cb[:, sk_c // 8:]
            pv = prev[5 + i] if prev is not None else None
            if self.precision == "bf16":
                if pv is not None:
                    # the up-conv GEMM has no `prev` input: keep this frame's hand-over slice, then splice in place
                    state[5 + i] = up[:, :1].clone()
                    call("uncl_splice_channels", up, st(up), pv, st(pv), up_c // 32, n, up_s * up_s, dt)
                call("uncl_convT2x2_tc", up, st(up), P["u%d_up" % i][0], P["u%d_up" % i][1], dst, st(cb), _lib.BF16, n,
                     up_c, up_s, up_s, sk_s, sk_s)
            elif self.precision == "fp32_tc":
                if pv is not None:
                    state[5 + i] = up[:, :1].clone()
                    call("uncl_splice_channels", up, st(up), pv, st(pv), up_c // 32, n, up_s * up_s, dt)
                xs = torch.empty((n, 3 * up_c // 8, up_s, up_s, 8), device=dev, dtype=torch.bfloat16)
                call("uncl_split_bf16", up, st(up), xs, st(xs), n, up_c, up_s * up_s)
                call("uncl_convT2x2_tc_cin", xs, st(xs), P["u%d_up" % i][0], P["u%d_up" % i][1], dst, st(cb), _lib.F32, n,
                     3 * up_c, up_c, up_s, up_s, sk_s, sk_s)
            else:
                call("uncl_convT2x2", up, st(up), pv, st(pv) if pv is not None else 0,
                     up_c // 32 if pv is not None else 0, P["u%d_up" % i][0], P["u%d_up" % i][1], dst, st(cb), n, up_c,
                     up_s, up_s, sk_s, sk_s, dt)
            co = f if i >= 2 else up_c // 2
            mid = buf(co, sk_s + 2, sk_s + 2)
            if 3 - i < nfused:
                rows = P.get("u%d_0_rows" % i) if self.row_kernel else None
                if rows is not None:
                    call("uncl_conv3x3_tc_rows_skipcat", cb, st(cb), rows, P["u%d_0" % i][0], P["u%d_0" % i][1], mid, st(mid), n,
                         sk_c, sk_s, sk_s, 2, ACT_RELU)
                else:
                    call("uncl_conv3x3_tc_skipcat", cb, st(cb), P["u%d_0" % i][0], P["u%d_0" % i][1], mid, st(mid), _lib.BF16, n,
                         sk_c, sk_s, sk_s, co, 2, ACT_RELU)
            else:
                self._conv3(P, "u%d_0" % i, cb, st(cb), mid, st(mid), n, 4 * sk_c, sk_s, sk_s, co, 2)
            last = i == 3
            if last and self.precision == "bf16" and not want_features:
                self._conv3(P, "u3_1", mid, st(mid), None, 0, n, co, sk_s + 2, sk_s + 2, co, 2,
                            fuse=(P["outc"][0], P["outc"][1], out, logit))
                fused, nxt = True, None
            else:
                nxt = buf(co, sk_s + 4, sk_s + 4)
                self._conv3(P, "u%d_1" % i, mid, st(mid), nxt, st(nxt), n, co, sk_s + 2, sk_s + 2, co, 2)
            up, up_c, up_s = nxt, co, sk_s + 4
            state.append(up)
            if keep is not None:
                keep.setdefault("ups", []).append(up)
        if not fused:
            call("uncl_outc_sigmoid", up, st(up), P["outc"][0], P["outc"][1], out, logit, n, f, 256 * 256, dt)
        return out, up, logit, state

    def _features_nchw(self, up):
        n = up.shape[0]
        c, hw = up.shape[1] * 8, up.shape[2] * up.shape[3]
        o = torch.empty((n, c, up.shape[2], up.shape[3]), device=up.device, dtype=torch.float32)
        call("uncl_blocked_to_nchw", up, up.stride(0), o, n, c, hw, _lib.DTYPE_OF[up.dtype])
        return o

    def _forward_train(self, x, droppath_scale=None, prev=None, want_state=False):
        """Same network as _run_frame, built from autograd Functions whose forward and backward are library kernels.
        precision 'fp32': CUDA-core kernels throughout (gradient parity 1e-3 vs the float64 oracle).
        precision 'bf16': "mixed" - the 3x3 convolutions (forward, data and weight gradients) run on the tcgen05
        kernels with bf16-rounded operands and fp32 accumulation; every tensor between kernels is fp32.
        prev / want_state: recurrent channel hand-over of the video generator (Unet.py:229-286)."""
        from . import autograd as A
        # 'fp32_tc': the 3x3 convolutions (98 % of the FLOPs; forward, data and weight gradients) as three-term bf16 split
        # GEMMs on the tensor cores (~2^-16 per product), everything else on the fp32 CUDA-core kernels
        tc = True if self.precision == "bf16" else ("split" if self.precision == "fp32_tc" else False)
        tc_other = tc is True   # up-convs and the graph block: tensor cores only on the mixed path
        if x.dim() != 4 or tuple(x.shape[1:]) != (1, 256, 256) or not x.is_cuda:
            raise ValueError("the generator expects CUDA [N,1,256,256] inputs")
        n = x.shape[0]
        c = self.inc.conv
        a0 = A.ConvFirst.apply(x, c.conv.weight, c.conv.bias)
        cur = A.Conv3x3.apply(a0, c.conv1.weight, c.conv1.bias, False, True, tc)
        skips, state = [cur], [cur]
        for i in range(4):
            blk = self.down_path[i].mpconv[1]
            fea = cur if prev is None else A.SpliceChannels.apply(cur, prev[i], cur.shape[1] * 8 // 32)
            m = A.Conv3x3.apply(A.MaxPool2.apply(fea), blk.conv.weight, blk.conv.bias, False, True, tc)
            cur = A.Conv3x3.apply(m, blk.conv1.weight, blk.conv1.bias, i == 3, True, tc)
            skips.append(cur)
            state.append(cur)
        g, ffn = self.gcn.module[0][0], self.gcn.module[0][1]
        s0 = droppath_scale[0] if droppath_scale is not None else None
        s1 = droppath_scale[1] if droppath_scale is not None else None
        x0 = A.AddPos.apply(cur, self.gcn.pos_embed)
        y = A.PwConv.apply(x0, g.fc1[0].weight, g.fc1[0].bias, None, None, 1, False, False)   # feeds KNN: fp32
        z = A.KnnAggregate.apply(y, g.relative_pos.detach().reshape(144, 144).float().contiguous())
        gc = g.graph_conv.gconv.nn[0]
        z2 = A.PwConv.apply(z, gc.weight, gc.bias, None, None, 4, True, tc_other)
        x1 = A.PwConv.apply(z2, g.fc2[0].weight, g.fc2[0].bias, x0, s0, 1, False, tc_other)
        f1 = A.PwConv.apply(x1, ffn.fc1[0].weight, ffn.fc1[0].bias, None, None, 1, True, tc_other)
        up = A.PwConv.apply(f1, ffn.fc2[0].weight, ffn.fc2[0].bias, x1, s1, 1, False, tc_other).reshape(n, -1, 12, 12, 8)
        state.append(up)
        for i in range(4):
            u = self.up_path[i]
            sk = skips[3 - i]
            fea = up if prev is None else A.SpliceChannels.apply(up, prev[5 + i], up.shape[1] * 8 // 32)
            x1u = A.ConvT2x2.apply(fea, u.up.weight, u.up.bias, sk.shape[2], sk.shape[3], tc_other)
            cat = A.SkipConcat.apply(sk, x1u)
            m = A.Conv3x3.apply(cat, u.conv.conv.weight, u.conv.conv.bias, True, True, tc)
            up = A.Conv3x3.apply(m, u.conv.conv1.weight, u.conv.conv1.bias, True, True, tc)
            state.append(up)
        if want_state:
            return A.OutcSigmoid.apply(up, self.outc.conv.weight, self.outc.conv.bias), A.BlockedToNCHW.apply(up), state
        out = A.OutcSigmoid.apply(up, self.outc.conv.weight, self.outc.conv.bias)
        return out, A.BlockedToNCHW.apply(up)

    def _droppath_scale(self, n, device):
        """Per-sample DropPath factors (mask / keep_prob) for the two residual branches; None in eval."""
        if not self.training or self.drop_path_prob <= 0:
            return None
        keep = 1.0 - self.drop_path_prob
        return [torch.empty(n, device=device).bernoulli_(keep) / keep for _ in range(2)]

    @staticmethod
    def _crop(x_out, diffY, diffX):
        # utils/data_loader_util.py:165-172 (crop_input_hdr_batch)
        h, w = x_out.shape[-2], x_out.shape[-1]
        th, tw = h - diffY, w - diffX
        i, j = int(round((h - th) / 2.0)), int(round((w - tw) / 2.0))
        return x_out[..., i:i + th, j:j + tw]


def blocked_to_nchw(t):
    """Debug/test helper: C8-blocked [N][C/8][H][W][8] (fp32 or bf16) -> NCHW fp32."""
    n, cb, h, w, _ = t.shape
    o = torch.empty((n, cb * 8, h, w), device=t.device, dtype=torch.float32)
    call("uncl_blocked_to_nchw", t, t.stride(0), o, n, cb * 8, h * w, _lib.DTYPE_OF[t.dtype])
    return o


class UNet(_GeneratorBase):
    """Image generator: forward(x[N,1,256,256]) -> (sigmoid map [N,1,256,256], up_x [N,32,256,256]).

    Drop-in for models/unet_multi_filters/Unet_singleFrame.py:101-213.  With autograd enabled the forward is built
    from `uncltmo_b200.autograd` Functions (fp32 path), so `loss.backward()` runs the library's backward kernels.
    """

    def forward_blocked(self, x):
        """(sigmoid map [N,1,256,256] fp32, up_x as the C8-blocked bf16 tensor [N,4,256,256,8] the network computed it in).
        The fast trainer's entry (uncltmo_b200.trainer): losses.infoNCE2 reads the blocked features directly, so the
        134 MB NCHW fp32 copy of `forward` and its gradient never exist.  bf16 precision only.  Without autograd the
        features are not materialised at all (second value None)."""
        if self.precision != "bf16":
            raise RuntimeError("forward_blocked is the bf16 path's entry")
        scale = self._droppath_scale(x.shape[0], x.device)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return self._train_bf16(x, scale)
        out, _, _, _ = self._run_frame(x, droppath_scale=scale, want_features=False)
        return out, None

    def _train_bf16(self, x, scale, through_autograd=False):
        """bf16-activation training pass as ONE autograd node (uncltmo_b200/train_graph.py).  through_autograd: parameter
        gradients are returned to autograd (hooks / DDP / any optimizer see them) instead of being accumulated into the flat
        gradient buffer as a side effect."""
        from .train_graph import GeneratorTrainFn, flat_params
        if x.dim() != 4 or tuple(x.shape[1:]) != (1, 256, 256) or not x.is_cuda:
            raise ValueError("the generator expects CUDA [N,1,256,256] inputs")
        fp = flat_params(self)
        # a fresh leaf per call: autograd ties a leaf's accumulation node to the stream it is first used on, and a
        # persistent one created in an eager step on the legacy stream would poison a later CUDA-graph capture
        anchor = torch.zeros(1, device=x.device, requires_grad=True)
        if through_autograd:
            return GeneratorTrainFn.apply(x, anchor, self, scale, *[p for _, p in fp.named if p.requires_grad])
        return GeneratorTrainFn.apply(x, anchor, self, scale)

    def forward(self, x, apply_crop=True, diffY=0, diffX=0):
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            if self.precision == "bf16" and not x.requires_grad:
                from . import autograd as A
                out, up = self._train_bf16(x, self._droppath_scale(x.shape[0], x.device), through_autograd=True)
                feats = A.BlockedToNCHW.apply(up)
            else:
                out, feats = self._forward_train(x, self._droppath_scale(x.shape[0], x.device))
        else:
            out, up, _, _ = self._run_frame(x, droppath_scale=self._droppath_scale(x.shape[0], x.device))
            feats = self._features_nchw(up)
        if apply_crop and self.to_crop:
            out = self._crop(out, diffY, diffX)
        return out, feats

    def tonemap_tiles(self, x, want_logit=False):
        """Fast path for the frame pipeline: [N,1,256,256] -> [N,1,256,256] without materialising features."""
        out, _, logit, _ = self._run_frame(x, want_features=False, want_logit=want_logit)
        return (out, logit) if want_logit else out


class UNetVideo(_GeneratorBase):
    """Video generator: forward(x[N,T,1,256,256]) -> (frames [N,T,1,256,256], features [N,T,64,1,1]).

    Drop-in for models/unet_multi_filters/Unet.py:135-289: frame k receives the first C/32 channels of its
    eight stage inputs from frame k-1.
    """

    def _train_clip_bf16(self, x, through_autograd):
        """bf16 training pass over a clip as one autograd node (train_graph.VideoTrainFn): ([N,T,1,256,256] frames,
        [N,T,64,1,1] features)."""
        from . import autograd as A
        from .features import plane_mean_contrast
        from .train_graph import VideoTrainFn, flat_params
        if x.dim() != 5 or tuple(x.shape[2:]) != (1, 256, 256) or not x.is_cuda:
            raise ValueError("the video generator expects CUDA [N,T,1,256,256] inputs")
        fp = flat_params(self)
        t_len = x.shape[1]
        scales = [self._droppath_scale(x.shape[0], x.device) for _ in range(t_len)]
        anchor = torch.zeros(1, device=x.device, requires_grad=True)
        params = [p for _, p in fp.named if p.requires_grad] if through_autograd else []
        res = VideoTrainFn.apply(x, anchor, self, scales if scales[0] is not None else None, *params)
        outs, feats = [], []
        for k in range(t_len):
            out, up = res[2 * k], res[2 * k + 1]
            mean, con = plane_mean_contrast(A.BlockedToNCHW.apply(up))
            feats.append(torch.cat([mean, con], dim=1)[:, None, :, None, None])
            outs.append(out.unsqueeze(1))
        return torch.cat(outs, 1), torch.cat(feats, 1)

    def forward_blocked(self, x):
        """The fast trainer's entry (uncltmo_b200.trainer): same values as forward(), parameter gradients accumulated
        straight into the flat gradient buffer (no per-parameter autograd work).  bf16 precision only."""
        if self.precision != "bf16":
            raise RuntimeError("forward_blocked is the bf16 path's entry")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return self._train_clip_bf16(x, through_autograd=False)
        return self.forward(x)

    def forward(self, x, apply_crop=True, diffY=0, diffX=0):
        from .features import contrast_features, plane_mean_contrast
        train = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        if train and self.precision == "bf16" and not x.requires_grad and not self.to_crop:
            return self._train_clip_bf16(x, through_autograd=True)
        outs, feats, prev = [], [], None
        for k in range(x.shape[1]):
            scale = self._droppath_scale(x.shape[0], x.device)
            if train:
                out, up_nchw, prev = self._forward_train(x[:, k].contiguous(), scale, prev=prev, want_state=True)
                mean, con = plane_mean_contrast(up_nchw)
                feats.append(torch.cat([mean, con], dim=1)[:, None, :, None, None])
            else:
                out, up, _, prev = self._run_frame(x[:, k], prev=prev, droppath_scale=scale)
                feats.append(contrast_features(up).unsqueeze(1))
            if apply_crop and self.to_crop:
                out = self._crop(out, diffY, diffX)
            outs.append(out.unsqueeze(1))
        return torch.cat(outs, 1), torch.cat(feats, 1)

    def tonemap_clip_tiles(self, frames):
        """frames: list over T of [N,1,256,256] tile batches (same tiles, consecutive frames) -> list of [N,1,256,256].
        No feature extraction (inference path of run_model_on_video)."""
        outs, prev = [], None
        for x in frames:
            out, _, _, prev = self._run_frame(x, prev=prev, want_features=True)
            outs.append(out)
        return outs
