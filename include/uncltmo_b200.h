/* uncltmo_b200 - C ABI of the B200-native UnCLTMO hot path (libuncltmo_b200.so, sm_100a only).
 *
 * The reference (cao-cong/UnCLTMO) has no FFI: its hot path sits behind Python classes that call stock
 * torch ops.  This header is the boundary a binding would use instead; each entry point names the
 * reference code it replaces.  Conventions:
 *   - every pointer is a DEVICE pointer unless stated; the caller allocates inputs, outputs and workspaces;
 *   - the library never allocates, never synchronises and launches only on `stream`;
 *   - return 0 on success or a negative UNCL_E* code; uncl_last_error() gives a thread-local message;
 *   - activations are "C8-blocked": [N][C/8][H][W][8] elements of `dtype` (0 = fp32, 1 = bf16); `*_img_stride`
 *     is the distance between images in ELEMENTS (so a tensor may be a channel-slice of a concat buffer);
 *   - plain images (network input / output, losses) are dense [N][H][W] fp32.
 */
#ifndef UNCLTMO_B200_H
#define UNCLTMO_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* uncl_stream_t; /* == cudaStream_t */

#define UNCL_OK 0
#define UNCL_EINVAL (-1)
#define UNCL_ECUDA (-2)
#define UNCL_EUNSUPPORTED (-3)

/* ---- library ---- */
const char* uncl_last_error(void);
const char* uncl_arch(void);   /* "sm_100a" */
int uncl_version(void);
int uncl_probe_device(int* out_dev, uncl_stream_t stream);   /* writes __CUDA_ARCH__ (+1 if arch-specific) */

/* ---- generator building blocks (models/unet_multi_filters/unet_parts.py) ---- */

/* inconv first conv: Conv2d(1, C_out, 3) + act.  unet_parts.py:57-87 (double_conv.conv), 196-203.
 * x [N][H][W] fp32; w [9][C_out] fp32; out blocked [N][C_out/8][H-2][W-2][8]. */
int uncl_conv_first(const float* x, const float* w, const float* bias, void* out, long out_img_stride, int N, int H,
                    int W, int C_out, int act, int dtype, uncl_stream_t stream);

/* 3x3 stride-1 conv, CUDA-core fp32-accumulate path.  pad 0 = nn.Conv2d valid (unet_parts.py:18,27);
 * pad 2 = nn.ConvTranspose2d(k=3, s=1, p=0) with flipped weights (unet_parts.py:114, 148-159).
 * w [9][C_in][C_out] fp32.  emit_skip=1 additionally writes y^2 and sqrt(y+1e-8) at channel-block offsets
 * 2*C_out/8 and 3*C_out/8 of `out` (concat operator 'square_and_square_root', unet_parts.py:319-322). */
int uncl_conv3x3_simt(const void* in, long in_img_stride, const float* w, const float* bias, void* out,
                      long out_img_stride, int N, int C_in, int H, int W, int C_out, int pad, int act, int emit_skip,
                      int dtype, uncl_stream_t stream);

/* Same operator on the tcgen05 tensor cores (bf16 operands, fp32 accumulation in TMEM, TMA-fed).
 * w_packed: bf16, the B-operand layout uncltmo_b200/packing.py:conv3x3_tc produces for this (C_in, C_out):
 * [NS][C_in/16][9][2][NT][8], or [NS][C_in/16][3][2][3*NT][8] for the kx-merged kernel (C_out <= 64, C_in >= 64); bias fp32.
 * fuse_outc=1: apply the 1x1 out conv (outc_w [C_out], outc_b [1]) + sigmoid in the epilogue and write
 * out_img [N][Ho][Wo] fp32 (unet_parts.py:338-345 + nn.Sigmoid, Unet_singleFrame.py:207-209); `out` may then be
 * NULL to skip storing the feature map.  The input is always bf16; `out_dtype` selects bf16 or fp32 stores (the
 * training path keeps fp32 tensors between the tensor-core convolutions).  The data gradient of a conv is this same
 * call on the output gradient with the taps reversed / transposed and pad 2 - pad. */
int uncl_conv3x3_tc(const void* in, long in_img_stride, const void* w_packed, const float* bias, void* out,
                    long out_img_stride, int out_dtype, int N, int C_in, int H, int W, int C_out, int pad, int act,
                    int emit_skip, int fuse_outc, const float* outc_w, const float* outc_b, float* out_img,
                    float* out_logit, uncl_stream_t stream);

/* Data gradient of the same operator with the ReLU backward of the PRODUCING layer fused into the epilogue:
 * out = corr(dZ, w_packed) * (mask > 0).  in = dZ bf16 blocked; w_packed = packing.conv3x3_tc of the flipped / transposed
 * taps; pad = 2 - (forward pad); mask = the forward INPUT of the conv (post-ReLU output of the previous layer, bf16 blocked,
 * C_out channels, image stride mask_img_stride elements; NULL = no mask).  out: bf16 or fp32 blocked - the previous
 * layer's pre-activation gradient, ready to be the operand of its own gradient GEMMs.  torch autograd of
 * unet_parts.py:57-87 / 183-193 (conv -> ReLU chains). */
int uncl_conv3x3_tc_dgrad(const void* in, long in_img_stride, const void* w_packed, const void* mask, long mask_img_stride,
                          void* out, long out_img_stride, int out_dtype, int N, int C_in, int H, int W, int C_out, int pad,
                          uncl_stream_t stream);

/* up.conv.conv with the skip operators fused: unet_parts.py:311-332 computes cat([x2, x1, x2^2, sqrt(x2 + 1e-8)]) and feeds
 * it to ConvTranspose2d 3x3.  Here `in` holds only [x2 (C_skip) | x1 (C_skip)] (bf16 blocked, 2*C_skip channels); the
 * squared and square-root planes are built in shared memory from the x2 chunk the kernel has just loaded and never touch
 * HBM.  w_packed = packing.conv3x3_tc of the full [9][4*C_skip][C_out] bank in concat order.  C_out == 32, C_skip % 32 == 0. */
int uncl_conv3x3_tc_skipcat(const void* in, long in_img_stride, const void* w_packed, const float* bias, void* out,
                            long out_img_stride, int out_dtype, int N, int C_skip, int H, int W, int C_out, int pad,
                            int act, uncl_stream_t stream);

/* Tile plan of uncl_conv3x3_tc_skipcat (the 16 fields of uncl_conv3x3_tc_plan): pure host arithmetic, no GPU needed. */
int uncl_conv3x3_tc_skipcat_plan(int N, int C_skip, int H, int W, int C_out, int pad, int* plan);

/* The narrow layers of the generator at 122..256 pixels (inc.conv1, down0.conv*, up2/up3.conv*: unet_parts.py:57-87,
 * :126-141, :183-193) through the row kernel (conv_tc_rows.cu): the three ky taps share one A read in N' = 3 C_out, an M
 * block is one image row of a 126-column band, every input row enters shared memory once per strip and the three terms
 * of an output pixel meet in one TMEM lane (added by the MMAs themselves in the ring variant, by the epilogue from three
 * accumulator slots otherwise).  Arguments as uncl_conv3x3_tc with C_out = 32 or 64 and bf16 blocked output (fuse_outc:
 * C_out = 32); w_rows = packing.conv3x3_tc_rows(w9), w_tail = packing.conv3x3_tc(w9) serves the output columns past the
 * last whole band (uncl_conv3x3_tc_rows_plan field 2; NULL allowed when that is 0). */
int uncl_conv3x3_tc_rows(const void* in, long in_img_stride, const void* w_rows, const void* w_tail, const float* bias,
                         void* out, long out_img_stride, int N, int C_in, int H, int W, int C_out, int pad, int act,
                         int emit_skip, int fuse_outc, const float* outc_w, const float* outc_b, float* out_img,
                         float* out_logit, uncl_stream_t stream);

/* inconv of the generator in one launch: Conv2d(1, 32, 3) + ReLU + Conv2d(32, 32, 3) + ReLU (unet_parts.py:57-87 with in_ch = 1,
 * Unet_singleFrame.py inc).  The first conv's 32-channel output never reaches memory: four extra warps of the row kernel
 * compute it into the stage ring (im2col rows in shared memory, three-term bf16 split MMAs, bias + ReLU) while the second
 * conv consumes it.  x: [N][H0][W0] fp32, image stride x_img_stride elements; fw = packing.conv_first_rows(w1, b1) (the
 * bias rides on a constant-one tap); w_rows = packing.conv3x3_tc_rows(w9), bias [32]; out: bf16 blocked, (H0-4) x (W0-4), emit_skip as in uncl_conv3x3_tc.
 * W0 - 4 must be a multiple of 126 or < 126 (252 for the generator's 256-pixel tiles). */
int uncl_conv_first_conv3x3_tc_rows(const float* x, long x_img_stride, const void* fw, const void* w_rows, const float* bias,
                                    void* out, long out_img_stride, int N, int H0, int W0, int act,
                                    int emit_skip, uncl_stream_t stream);

/* uncl_conv3x3_tc_skipcat (unet_parts.py:311-332, skip operators built in shared memory, C_out = 32) through the row kernel.
 * `in` = [x2 (C_skip) | x1 (C_skip)] bf16 blocked; w_rows / w_tail = packing.conv3x3_tc_rows / packing.conv3x3_tc of the
 * full [9][4*C_skip][32] bank in concat order.  4*C_skip*576 B of filters must fit shared memory (C_skip = 32, 64). */
int uncl_conv3x3_tc_rows_skipcat(const void* in, long in_img_stride, const void* w_rows, const void* w_tail,
                                 const float* bias, void* out, long out_img_stride, int N, int C_skip, int H, int W, int pad,
                                 int act, uncl_stream_t stream);

/* Plan of the row kernel: pure host arithmetic, no GPU needed.  C_in = logical input channels (4*C_skip when derive = 1),
 * sms = SMs to balance the strips over.  plan[16] = { eligible (0: the filter bank does not fit - use uncl_conv3x3_tc),
 *   columns computed by the row kernel, trailing columns left to the older kernels, bands, band width, strip height,
 *   strips per band, work items, rows per pipeline stage, K chunks per row group, pipeline stages, stage bytes, resident
 *   filter bytes, dynamic shared memory bytes, sms, ring variant (1) or slot variant (0) of the accumulators }. */
int uncl_conv3x3_tc_rows_plan(int N, int C_in, int H, int W, int C_out, int pad, int derive, int sms, int* plan);

/* Tile plan uncl_conv3x3_tc would use for a problem: pure host arithmetic, callable without a GPU (tests check the tile
 * coverage and that uncltmo_b200/packing.py packs the weights for the kernel the library will pick).
 * plan[16] = { kind (0: one tap per MMA, conv_tc.cu; bit 0: kx-merged, conv_tc_merged.cu, bit 1: row-aligned tiles, bit 2: resident weights), NT, NS, MMA N, M blocks per tile,
 *   tile advance in positions, PW, PH, BW, column bands, tiles per band, work items, pipeline stages, accumulator stages,
 *   K=16 steps per stage, dynamic shared memory bytes }. */
int uncl_conv3x3_tc_plan(int N, int C_in, int H, int W, int C_out, int pad, int* plan);

/* up.up: nn.ConvTranspose2d(C, C, 2, stride=2) + bias, written into an (H2 x W2) channel slice of the skip
 * concat buffer with F.pad(..., mode='replicate') semantics.  unet_parts.py:283-299.  w [C][4][C] fp32.
 * prev/r: video generator recurrence - the first r input channels come from `prev` (Unet.py:270). */
int uncl_convT2x2(const void* in, long in_img_stride, const void* prev, long prev_img_stride, int r, const float* w,
                  const float* bias, void* out, long out_img_stride, int N, int C, int H, int W, int H2, int W2,
                  int dtype, uncl_stream_t stream);

/* The same ConvTranspose k2 s2 as a tcgen05 GEMM [pixels x C].[C x 4C] with a pixel-shuffle epilogue (bf16).
 * w_packed: bf16 [NS][C/16][1][2][NT][8], column j = (dy*2+dx)*C + co, NT = min(4C, 128) (packing.convT2x2_tc). */
int uncl_convT2x2_tc(const void* in, long in_img_stride, const void* w_packed, const float* bias, void* out,
                     long out_img_stride, int out_dtype, int N, int C, int H, int W, int H2, int W2,
                     uncl_stream_t stream);

/* The same GEMM with C_in != C: the exact path's three-term bf16 split, in = [x_hi | x_hi | x_lo] (C_in = 3C) from
 * uncl_split_bf16, w_packed = packing.convT2x2_tc_split ([w_hi ; w_lo ; w_hi]). */
int uncl_convT2x2_tc_cin(const void* in, long in_img_stride, const void* w_packed, const float* bias, void* out,
                         long out_img_stride, int out_dtype, int N, int C_in, int C, int H, int W, int H2, int W2,
                         uncl_stream_t stream);

/* Video recurrence for kernels without a `prev` input: dst[:, :r] = src[:, :r] (r <= 8).  Unet.py:244, 270. */
int uncl_splice_channels(void* dst, long dst_img_stride, const void* src, long src_img_stride, int r, int N, int HW,
                         int dtype, uncl_stream_t stream);

/* down.mpconv[0]: nn.MaxPool2d(2).  unet_parts.py:210-213.  prev/r as above (Unet.py:244). */
int uncl_maxpool2(const void* in, long in_img_stride, const void* prev, long prev_img_stride, int r, void* out,
                  long out_img_stride, int N, int C, int H, int W, int dtype, uncl_stream_t stream);

/* outconv 1x1 (C -> 1) + nn.Sigmoid.  unet_parts.py:338-345, Unet_singleFrame.py:207-209.
 * out / logit: [N][HW] fp32 (logit may be NULL). */
int uncl_outc_sigmoid(const void* in, long in_img_stride, const float* w, const float* b, float* out, float* logit,
                      int N, int C, int HW, int dtype, uncl_stream_t stream);

/* fp32 blocked [N][C/8][HW][8] (image stride in elements) -> bf16 blocked [N][3C/8][HW][8] = [hi | hi | lo] with
 * hi = bf16(x), lo = bf16(x - hi): the activation operand of the exact tensor-core path, where every product is the
 * three-term sum x_hi.w_hi + x_hi.w_lo + x_lo.w_hi accumulated in fp32 (relative error ~2^-16 instead of bf16's 2^-9). */
int uncl_split_bf16(const float* in, long in_img_stride, void* out, long out_img_stride, int N, int C, long HW,
                    uncl_stream_t stream);
/* dense fp32 <-> bf16 conversion of a blocked tensor (n elements, multiple of 8) */
int uncl_convert(const void* in, int in_dtype, void* out, int out_dtype, long n, uncl_stream_t stream);

/* layout transforms at the module boundary (NCHW fp32 <-> blocked) */
int uncl_blocked_to_nchw(const void* in, long in_img_stride, float* out, int N, int C, int HW, int dtype,
                         uncl_stream_t stream);
int uncl_nchw_to_blocked(const float* in, void* out, long out_img_stride, int N, int C, int HW, int dtype,
                         uncl_stream_t stream);

/* ---- bottleneck graph block (Unet_singleFrame.py:44-99, gcn_lib/) - fp32 blocked tensors, HW = 144 ---- */

/* x + pos_embed.  Unet_singleFrame.py:94.  pos: blocked fp32 [C/8][144][8]. */
int uncl_gcn_add_pos(const void* in, long in_img_stride, const float* pos, float* out, int N, int C, int dtype,
                     uncl_stream_t stream);

/* The same, additionally writing x0 as bf16 blocked [N][3C/8][144][8] = [hi | hi | lo] (hi = bf16(x0), lo = bf16(x0 - hi)):
 * the A operand of the three-term bf16 GEMM that computes Grapher fc1 (torch_vertex.py:219) on the tensor cores to
 * ~2^-16 relative accuracy; weights from uncltmo_b200/packing.py:pointwise_tc_split, GEMM = uncl_pw_conv_tc. */
int uncl_gcn_add_pos_split(const void* in, long in_img_stride, const float* pos, float* out, void* split, int N, int C,
                           int dtype, uncl_stream_t stream);

/* 1x1 conv (+groups, bias, act, residual, per-sample DropPath scale): Grapher fc1/fc2, BasicConv (groups=4), FFN.
 * gcn_lib/torch_vertex.py:219-227, gcn_lib/torch_nn.py:54-78, Unet_singleFrame.py:36-42.
 * w [groups][C_in/groups][C_out/groups] fp32; out = scale[n] * act(conv + bias) + res. */
int uncl_pw_conv(const float* in, const float* w, const float* bias, const float* res, const float* scale, void* out,
                 long out_img_stride, int N, int C_in, int C_out, int groups, int HW, int act, int out_dtype,
                 uncl_stream_t stream);

/* DenseDilatedKnnGraph (k=9, dilation 1, relative_pos) + MRConv2d aggregation + channel interleave.
 * gcn_lib/torch_edge.py:135-159, 54-86, 9-20; gcn_lib/torch_vertex.py:21-30.
 * y [N][C/8][144][8]; relpos [144][144]; z [N][2C/8][144][8] (fp32 or bf16); idx_out [N][144][9] int32 or NULL. */
int uncl_gcn_knn_aggregate(const float* y, const float* relpos, void* z, int z_dtype, int* idx_out, int N, int C,
                           uncl_stream_t stream);

/* Diagnostics for profiling only: per-role cycle counters of the tensor-core conv kernels (12 uint64 on the device,
 * zeroed by the caller; NULL switches off).  See conv_tc.cu. */
int uncl_conv_tc_set_debug(void* counters);

/* The same 1x1 (grouped) conv as a tcgen05 GEMM per group, bf16 operands / fp32 accumulation:
 * out = scale[n] * act(W x + b) + res, act in {none, ReLU, GELU}.  Also the data gradient of the k2 s2 up-convolution
 * (4C -> C over the space-to-depth gradient).  in: bf16 blocked [N][C_in/8][H][W][8], W <= 128;
 * w_packed: bf16 [NS][C_in/g/16][2][NT][8], NT = min(C_out/g, 128) (packing.pointwise_tc); bias / res / scale may be
 * NULL; res and out are blocked with C_out channels, fp32 or bf16 each. */
int uncl_pw_conv_tc(const void* in, long in_img_stride, const void* w_packed, const float* bias, const void* res,
                    long res_img_stride, int res_dtype, const float* scale, void* out, long out_img_stride,
                    int out_dtype, int N, int C_in, int C_out, int groups, int H, int W, int act,
                    uncl_stream_t stream);

/* Pointwise data gradient with a fused ReLU mask (see uncl_conv3x3_tc_dgrad): dX of the k2 s2 up-convolution. */
int uncl_pw_conv_tc_dgrad(const void* in, long in_img_stride, const void* w_packed, const void* mask, long mask_img_stride,
                          void* out, long out_img_stride, int out_dtype, int N, int C_in, int C_out, int groups, int H, int W,
                          uncl_stream_t stream);

/* ---- frame path (utils/model_save_util.py, utils/hdr_image_util.py, utils/data_loader_util.py) ---- */

/* bytes of the scratch buffer the frame-path calls below share (statistics, percentile state, histograms) */
long uncl_frame_workspace_bytes(void);

/* log-lambda normalisation fused with the replicate pad:  Y = .299R+.587G+.114B (after rgb -= min(rgb) if negative),
 * Y -= min; out = log10(Y/max*f + 1) / max(...), written into the (H1 x W1) padded frame.
 * model_save_util.py:232-239 (load_inference2), 255-262; hdr_image_util.py:76-82; data_loader_util.py:135-157, 175-179.
 * rgb [3][H][W] fp32.  stats_out[4] receives (min rgb, min Y, max Y, -) for uncl_frame_postprocess. */
int uncl_frame_normalise_pad(const float* rgb, int H, int W, float f_factor, float* gray_out, int H1, int W1,
                             float* stats_out, void* workspace, uncl_stream_t stream);

/* gather T 256x256 tiles at origins[t] = (y, x) from the padded frame.  model_save_util.py:417-426, 438, 455-459, 470. */
int uncl_tiles_gather(const float* frame, int H1, int W1, const int* origins, int T, float* tiles,
                      uncl_stream_t stream);

/* closed form of the sequential linear cross-fade of model_save_util.py:428-481: every pixel is a weighted sum of
 * at most KxK tiles (K=3 at overlap 64); per-row / per-column [L][K] (tile index, weight) tables are built by the
 * host mirror (uncltmo_b200/frame.py). */
int uncl_tiles_blend(const float* tiles, const int* yidx, const float* yw, const int* ystart, const int* xidx,
                     const float* xw, const int* xstart, int TX, int K, float* out, int H1, int W1,
                     uncl_stream_t stream);

/* numpy.percentile(clip(data, clamp_lo, clamp_hi), [p_lo, p_hi]) ('linear'), on device, no host sync.
 * model_save_util.py:389-390 (99.5 / 0.5), hdr_image_util.py:93-102 (99.0 / 0.1).  pct_out[2]. */
int uncl_percentile_pair(const float* data, long n, float clamp_lo, float clamp_hi, double p_lo, double p_hi,
                         float* pct_out, void* workspace, uncl_stream_t stream);

/* clamp to pct, min-max stretch, (rgb/(Y+1e-8))^0.5 * fake, crop the pad.  model_save_util.py:393-402,
 * hdr_image_util.py:122-132.  fake [H1][W1]; rgb / out [3][H][W]; stats from uncl_frame_normalise_pad. */
int uncl_frame_postprocess(const float* fake, int H1, int W1, const float* rgb, int H, int W, const float* stats,
                           const float* pct, float* out, uncl_stream_t stream);

/* ---- the same frame stages fused into ONE cooperative launch each (bit-identical values; a 1080p frame is a handful of
 * 4-9 us memory passes, so the staged path above is bound by its 14 launches, not by HBM) ---- */

/* 1 if the three fused calls below can run for this geometry on the current device, else 0 (use the staged calls). */
int uncl_frame_fused_supported(int H, int W, int H1, int W1, int K);

/* uncl_frame_normalise_pad + uncl_tiles_gather: statistics, (shifted statistics), normalisation written directly as the
 * T 256x256 tiles at origins[t] = (y, x) of the padded frame - which is never stored.  model_save_util.py:232-239, 417-470. */
int uncl_frame_normalise_tiles(const float* rgb, int H, int W, float f_factor, int H1, int W1, const int* origins, int T,
                               float* tiles, float* stats_out, void* workspace, uncl_stream_t stream);

/* uncl_tiles_blend + uncl_percentile_pair of the blended plane (K <= 3): plane_out [H1][W1], pct_out[2].
 * model_save_util.py:428-481, 389-390. */
int uncl_frame_blend_percentiles(const float* tiles, const int* yidx, const float* yw, const int* ystart, const int* xidx,
                                 const float* xw, const int* xstart, int TX, int K, float* plane_out, int H1, int W1,
                                 double p_lo, double p_hi, float* pct_out, void* workspace, uncl_stream_t stream);

/* uncl_frame_postprocess + uncl_percentile_pair(clip(col,0,1)) + uncl_frame_to_u8: col_out [3][H][W] (may be NULL),
 * pct_out[2] = the colour percentiles, u8_out HWC.  model_save_util.py:393-402, hdr_image_util.py:93-102, 122-132, 237-245. */
int uncl_frame_post_u8(const float* fake, int H1, int W1, const float* rgb, int H, int W, const float* stats,
                       const float* pct_plane, float* col_out, double p_lo, double p_hi, float* pct_out,
                       unsigned char* u8_out, void* workspace, uncl_stream_t stream);

/* clamp(0,1), stretch between pct[0..1], clip, *255 -> uint8 HWC.  hdr_image_util.py:237-245, 93-102. */
int uncl_frame_to_u8(const float* col, int H, int W, const float* pct, unsigned char* out, uncl_stream_t stream);

/* ---- discriminator and training losses, forward (dense fp32 planes [M][H][W]) ---- */

/* per-plane mean and mean of the local variance under the 11x11 gaussian (sigma 1.5, valid):
 * ContrastExtracter / compute_contrast + adaptive_avg_pool2d.  models/Discriminator.py:49-83, 121-125;
 * Unet.py:101-133, 274-278; GanTrainerImg.py:24-56, 308-313.  scratch: 2*M floats.  Outputs may be NULL. */
int uncl_plane_mean_contrast(const float* x, long plane_stride, int M, int H, int W, float* mean_out,
                             float* cmean_out, float* scratch, uncl_stream_t stream);

/* F.interpolate(scale_factor=0.5, mode='bicubic', align_corners=False).  models/struct_loss.py:52-53. */
int uncl_bicubic_half(const float* in, float* out, int M, int H, int W, uncl_stream_t stream);

/* StructLoss.forward: sum_l w[l] * MSE(z-normalised 5x5 windows of fake_l, of hdr_l), pyramid by bicubic x0.5.
 * models/struct_loss.py:23-104.  weights_host: HOST array [levels].  loss_out: 1 float (device).
 * scratch: >= 2*M*(H/2)*(W/2)*4/3 floats. */
int uncl_struct_loss_fwd(const float* fake, const float* hdr, int M, int H, int W, int levels,
                         const float* weights_host, float* loss_out, float* scratch, uncl_stream_t stream);

/* SimpleDiscriminator.forward up to (fea map, logits): models/Discriminator.py:98-122.
 * w1 [16][1][4][4], w2 [32][16][4][4], w3 [32], w_tail [62*62]; h_scratch N*16*127*127 floats (the post-LeakyReLU
 * first activation, kept for the backward); a2_out N*32*62*62 floats or NULL (second activation, for the backward);
 * fea [N][62][62]; logits [N]. */
int uncl_disc_forward(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                      const float* w3, const float* b3, const float* w_tail, float* h_scratch, float* a2_out,
                      float* fea, float* logits, int N, int H, int W, uncl_stream_t stream);

/* GanTrainer.contrastive_D_loss.  GanTrainerImg.py:219-229. */
int uncl_contrastive_d_loss(const float* real_logits, const float* fake_logits, int B, float* out,
                            uncl_stream_t stream);

/* GanTrainer.nce with one positive and one negative, 'InfoNCE'.  GanTrainerImg.py:410-439.
 * anchor [B][C][HW]; pos / neg likewise with image stride pos_stride / neg_stride in elements (0 = one sample
 * broadcast, as infoNCE2 does :400-401).  logits_scratch: 2*B floats. */
int uncl_nce_fwd(const float* anchor, const float* pos, long pos_stride, const float* neg, long neg_stride, int B,
                 int C, int HW, float k, float constant, float* logits_scratch, float* loss_out,
                 uncl_stream_t stream);

/* TMQI statistical naturalness N of M images in [0,1] (TMQI.py:210-242, `original` block mode): the score that picks
 * the positives / negatives of infoNCE2 and the pseudo label (GanTrainerImg.py:341-408).  scratch: 2*M floats. */
int uncl_tmqi_naturalness(const float* x, int M, int H, int W, float* scratch, float* out, uncl_stream_t stream);

/* nn.L1Loss of two vectors (per-image means).  GanTrainerImg.py:308-313. */
int uncl_l1_mean(const float* a, const float* b, int n, float* out, uncl_stream_t stream);

/* L_TV.  GanTrainer.py:669-682.  scratch: 2 floats. */
int uncl_tv_loss(const float* x, int B, int C, int H, int W, float* scratch, float* out, uncl_stream_t stream);

/* ---- backward building blocks of the generator (fp32, C8-blocked; dense tensors unless a stride is given) ----
 * Data gradients of the 3x3 convs / pointwise convs reuse uncl_conv3x3_simt / uncl_pw_conv with transposed weights. */

/* dY *= (Y > 0) in place (apply_relu) and db[c] += sum dY (db may be NULL).  ReLU of unet_parts.py:70-86. */
int uncl_relu_bwd_bias(float* dY, const float* Y, long y_img_stride, float* db, int N, int C, int HW, int apply_relu,
                       uncl_stream_t stream);
/* Out-of-place form: dZ = dY * (Y > 0) as fp32 or bf16 (the tensor-core gradient operand), db[c] += sum dZ. */
int uncl_relu_bwd_bias_out(const float* dY, const float* Y, long y_img_stride, void* dZ, int dz_dtype, float* db, int N,
                           int C, int HW, int apply_relu, uncl_stream_t stream);
/* dW9[t][ci][co] += sum X[.., y+ky-pad, x+kx-pad, ci] * dZ[.., y, x, co]   (dW9 zeroed by the caller) */
int uncl_conv3x3_wgrad(const float* X, long x_img_stride, const float* dZ, float* dW9, int N, int C_in, int H, int W,
                       int C_out, int pad, uncl_stream_t stream);
/* The same weight gradient on the tcgen05 tensor cores (bf16 X and dZ, fp32 accumulation in TMEM, fp32 atomics into
 * dW9): a GEMM with K = pixels whose operands are read MN-major straight from the C8-blocked TMA tiles. */
int uncl_conv3x3_wgrad_tc(const void* X, long x_img_stride, const void* dZ, float* dW9, int N, int C_in, int H, int W,
                          int C_out, int pad, uncl_stream_t stream);
/* The same with an explicit dZ image stride (elements): dZ may be a channel slice of a wider tensor.  The exact training
 * path (precision 'fp32_tc') accumulates x_hi.dz_hi + x_hi.dz_lo + x_lo.dz_hi into one dW9 with three calls. */
int uncl_conv3x3_wgrad_tc_strided(const void* X, long x_img_stride, const void* dZ, long dz_img_stride, float* dW9, int N,
                                  int C_in, int H, int W, int C_out, int pad, uncl_stream_t stream);
/* Pointwise (GEMM) weight gradient on the same tensor-core kernel (one tap): dW[ci][co] += sum_pix X[pix,ci] * dZ[pix,co];
 * bf16 blocked operands with their own image strides, dW fp32 [C_in][C_out] (zeroed by the caller).  Weight gradient of
 * the k2 s2 up-convolution over the space-to-depth output gradient (C_out = 4C). */
int uncl_pw_wgrad_tc(const void* X, long x_img_stride, const void* dZ, long dz_img_stride, float* dW, int N, int C_in,
                     int C_out, int H, int W, uncl_stream_t stream);
int uncl_conv_first_wgrad(const float* x, const float* dZ, float* dW, int N, int H, int W, int C, uncl_stream_t stream);
int uncl_maxpool2_bwd(const float* X, long x_img_stride, const float* dP, float* dX, int N, int C, int H, int W,
                      uncl_stream_t stream);
/* cat = [x2 | x1 | x2^2 | sqrt(x2+1e-8)] (unet_parts.py:319-322) and its backward */
int uncl_skip_concat_fwd(const float* x2, long x2_img_stride, const float* x1, float* cat, int N, int C, int HW,
                         uncl_stream_t stream);
int uncl_skip_concat_bwd(const float* dcat, const float* x2, long x2_img_stride, float* dx2, float* dx1, int N, int C,
                         int HW, uncl_stream_t stream);
/* ConvTranspose k2 s2 backward helper: fold the replicate pad and move (dy,dx) into channels: [N][4C/8][H][W][8];
 * out_bf16 (may be NULL): a bf16 copy of the same tensor, the operand of the tensor-core data-gradient GEMM. */
int uncl_convT2x2_s2d(const float* dY, float* out, void* out_bf16, int N, int C, int H, int W, int H2, int W2,
                      uncl_stream_t stream);
/* dW[g][ci][co] += sum_pix X[pix,ci] * dZ[pix,co]   (dW zeroed by the caller) */
int uncl_pw_wgrad(const float* X, const float* dZ, float* dW, int N, int C_in, int C_out, int groups, int HW,
                  uncl_stream_t stream);
int uncl_gelu_fwd(const float* u, float* g, long n, uncl_stream_t stream);
int uncl_gelu_bwd(const float* u, const float* dg, float* du, long n, uncl_stream_t stream);
int uncl_scale_rows(float* x, const float* scale, int N, long per_image, uncl_stream_t stream);
int uncl_batch_sum(const float* x, float* out, int N, long M, uncl_stream_t stream);
/* MRConv2d aggregation backward (gcn_lib/torch_vertex.py:21-30); dy zeroed by the caller */
int uncl_gcn_agg_bwd(const float* dz, const float* y, const int* idx, float* dy, int N, int C, uncl_stream_t stream);
/* outconv + sigmoid backward; dw / db zeroed by the caller */
int uncl_outc_sigmoid_bwd(const float* d_out, const float* out, const float* up, long up_img_stride, const float* w,
                          float* d_up, float* dw, float* db, int N, int C, int HW, uncl_stream_t stream);

/* ---- bf16-activation training path (uncltmo_b200/train_graph.py): every activation and pre-activation gradient of the
 * generator is a C8-blocked bf16 tensor; these are the places where gradient paths meet or a layout changes ---- */

/* Gradient of a skip tensor x2 (post-ReLU output of an encoder conv; lives in its concat buffer, image stride
 * x2_img_stride): dz = (x2 > 0) * (dcat[0:C] + 2 x2 dcat[2C:3C] + 0.5 dcat[3C:4C] / sqrt(x2 + 1e-8) + maxpool2_bwd(dpool)),
 * i.e. the backward of cat([x2, x1, x2^2, sqrt(x2+1e-8)]) (unet_parts.py:319-322), of MaxPool2d(2) (:210-213; first
 * maximum in row-major order takes the gradient) and of the ReLU of the layer that produced x2, in one pass.
 * dcat: bf16 dense [N][4C/8][H][W][8] or NULL; dpool: bf16 dense [N][C/8][H/2][W/2][8] or NULL; dz: bf16 dense;
 * db[C] += column sums of dz (NULL to skip). */
int uncl_skip_pool_bwd(const void* x2, long x2_img_stride, const void* dcat, const void* dpool, void* dz, float* db, int N,
                       int C, int H, int W, uncl_stream_t stream);
/* The same with the video generator's recurrence (Unet.py:244: the pool read cat(prev[:, :r], x2[:, r:]), r <= 8 channels of
 * channel block 0): window values of channels < r come from `prev` (bf16 blocked, first block used), their pool gradient is
 * written to d_prev (bf16 dense [N][1][H][W][8], channels >= r zero) instead of dz; d_state (bf16 dense [N][1][H][W][8] or
 * NULL) - the gradient the NEXT frame sent to this frame's own first r channels - is added before the ReLU mask.
 * prev / d_prev both NULL: no splice at this level (first frame of a clip). */
int uncl_skip_pool_bwd_rec(const void* x2, long x2_img_stride, const void* dcat, const void* dpool, void* dz, float* db, int N,
                           int C, int H, int W, const void* prev, long prev_img_stride, int r, void* d_prev,
                           const void* d_state, uncl_stream_t stream);
/* Recurrence fix-up on a decoder tensor's gradient (Unet.py:270), channel block 0 of dz [N][C/8][HW][8] in place:
 * d_prev != NULL (the up-conv read prev's first r channels): d_prev[ch<r] = dz[ch], dz[ch<r] = 0;
 * d_state != NULL: dz[ch<r] += (own[ch] > 0) * d_state[ch] (own NULL: no mask). */
int uncl_splice_grad(void* dz, long dz_img_stride, const void* own, long own_img_stride, int r, void* d_prev,
                     const void* d_state, int N, long HW, uncl_stream_t stream);
/* db[c] += sum over images and pixels of a bf16 blocked tensor (bias gradient of a conv whose dz a GEMM epilogue wrote) */
int uncl_bias_grad_bf16(const void* dz, long img_stride, float* db, int N, int C, int HW, uncl_stream_t stream);
/* bf16 form of uncl_convT2x2_s2d reading a channel slice of the concat gradient (image stride dy_img_stride);
 * db[C] += the up-convolution's bias gradient (NULL to skip). */
int uncl_convT2x2_s2d_bf16(const void* dY, long dy_img_stride, void* out, float* db, int N, int C, int H, int W, int H2,
                           int W2, uncl_stream_t stream);
/* Weight AND bias gradient of the first conv (Conv2d(1, C, 3), unet_parts.py:57-87) in one pass over its pre-activation
 * gradient dZ (blocked dense [N][C/8][H-2][W-2][8], fp32 or bf16): dW[9][C] += ..., db[C] += ... (db may be NULL). */
int uncl_conv_first_wgrad_bias(const float* x, const void* dZ, int dz_dtype, float* dW, float* db, int N, int H, int W,
                               int C, uncl_stream_t stream);
/* 1x1 out conv + sigmoid backward (unet_parts.py:338-345, Unet_singleFrame.py:207-209) merged with the gradient that
 * arrives through the feature output and the ReLU of up_path.3.conv.conv1:  dl = d_out * o * (1 - o);
 * dz[c] = up[c] > 0 ? dl * w[c] + d_feat[c] : 0 (bf16 dense);  dw[c] += sum dl * up[c];  db_out += sum dl;
 * db_up[c] += sum dz[c].  d_out (fp32 [N][HW]) and d_feat (bf16 blocked dense) may each be NULL.  C = 32. */
int uncl_outc_feat_bwd(const float* d_out, const float* out, const void* up, long up_img_stride, const void* d_feat,
                       const float* w, void* dz, float* dw, float* db_out, float* db_up, int N, int C, int HW,
                       uncl_stream_t stream);
/* dst_bf16[i] = bf16(src[idx[i]]); idx bit 30 set: the bf16 residual src - bf16(src) instead (`lo` half of a split-bf16
 * operand); idx < 0: zero.  One launch re-lays out every weight of the generator for the forward and the data-gradient
 * GEMMs (index maps from uncltmo_b200/packing.py, built once). */
int uncl_pack_gather(const float* src, const int* idx, void* dst_bf16, long n, uncl_stream_t stream);
/* dst[i] = src[idx[i]] (idx < 0: dst untouched): fp32 gather, e.g. GEMM-layout weight gradients -> parameter layout. */
int uncl_unpack_gather(const float* src, const int* idx, float* dst, long n, uncl_stream_t stream);
/* dst[i] += src[idx[i]]: accumulating form (all GEMM-layout weight gradients of a backward pass into the flat p.grad). */
int uncl_unpack_add(const float* src, const int* idx, float* dst, long n, uncl_stream_t stream);
/* GanTrainer.nce as infoNCE2 calls it (GanTrainerImg.py:384-439): positive / negative are rows sel[0] / sel[1] (device
 * int64) of the anchor tensor itself, broadcast over the batch.  fea: bf16 [B][CHW] in ANY element order (the
 * similarity is a sum over all elements).  A row index >= B selects row (index - B) of `ext` (bf16 [2][CHW]) instead:
 * data-parallel training, where the global arg-max / arg-min sample may live on another rank (ext may be NULL when both
 * indices are < B).  logits_scratch: 2*B floats (kept for the backward). */
int uncl_nce_self_fwd(const void* fea, const long* sel, const void* ext, int B, long CHW, int HW, float k, float constant,
                      float* logits_scratch, float* loss_out, uncl_stream_t stream);
/* d_fea (bf16 or fp32, d_dtype): anchor gradient of every row + the batch-summed gradient of the selected rows that are
 * local; d_ext (fp32 [2][CHW], may be NULL without ext rows): this rank's partial gradient of the external rows. */
int uncl_nce_self_bwd(const void* fea, const long* sel, const void* ext, int B, long CHW, int HW, float k, float constant,
                      const float* logits, const float* g_up, void* d_fea, int d_dtype, float* d_ext,
                      uncl_stream_t stream);
/* torch.optim.Adam's update (no amsgrad / weight decay) on one flat fp32 buffer; `step` is a device float, incremented
 * by the call (bias corrections need no host value: capturable in a CUDA graph). */
int uncl_adam_flat(float* p, const float* g, float* m, float* v, long n, float lr, float beta1, float beta2, float eps,
                   float* step, uncl_stream_t stream);

/* ---- loss / discriminator backward.  `g_up` is the upstream gradient as a DEVICE scalar (no host sync). ---- */

/* d StructLoss / d fake.  scratch: >= 6.5*M*H*W + 64 floats.  weights_host: HOST array [levels]. */
int uncl_struct_loss_bwd(const float* fake, const float* hdr, int M, int H, int W, int levels,
                         const float* weights_host, const float* g_up, float* d_fake, float* scratch,
                         uncl_stream_t stream);
/* logits: the 2*B similarities uncl_nce_fwd left in its scratch.  d_pos / d_neg may be NULL. */
int uncl_nce_bwd(const float* anchor, const float* pos, long pos_stride, const float* neg, long neg_stride, int B,
                 int C, int HW, float k, float constant, const float* logits, const float* g_up, float* d_anchor,
                 float* d_pos, float* d_neg, uncl_stream_t stream);
int uncl_contrastive_d_bwd(const float* real_logits, const float* fake_logits, int B, const float* g_up, float* d_real,
                           float* d_fake, uncl_stream_t stream);
/* d_mean / d_cmean: [M] upstream gradients of uncl_plane_mean_contrast's outputs (either may be NULL);
 * mu_scratch: M*(H-10)*(W-10) floats. */
int uncl_plane_mean_contrast_bwd(const float* x, int M, int H, int W, const float* d_mean, const float* d_cmean,
                                 float* dx, float* mu_scratch, uncl_stream_t stream);
int uncl_l1_mean_bwd(const float* a, const float* b, int n, const float* g_up, float* da, float* db,
                     uncl_stream_t stream);
int uncl_tv_bwd(const float* x, int B, int C, int H, int W, const float* g_up, float* dx, uncl_stream_t stream);
/* SimpleDiscriminator backward (models/Discriminator.py:98-126).  d_fea [N][62][62] holds the gradient arriving
 * through the feature branch on entry (zeros if none); the tail's contribution is added.  dx may be NULL.
 * dw3 / db3 must be zeroed by the caller.  dw1 / db1 and dw2 / db2 may be NULL: the two convolutions' weight gradients are
 * then skipped (back-propagation THROUGH the discriminator in the generator step).  scratch: N*(32*62*62 + 16*127*127) floats. */
int uncl_disc_backward(const float* x, const float* h1, const float* a2, const float* fea, const float* w1,
                       const float* w2, const float* w3, const float* w_tail, const float* d_logits, float* d_fea,
                       float* dx, float* dw1, float* db1, float* dw2, float* db2, float* dw3, float* db3,
                       float* dw_tail, float* scratch, int N, uncl_stream_t stream);

/* ---- HDR file decode at the head of the entry path (utils/hdr_image_util.py:35-53 reads on the host) ---- */

/* HOST function (host pointers, no CUDA call): byte offset of every scanline of a Radiance RGBE (.hdr) file's pixel stream.
 * data_host: the whole file; pixel_offset: first byte after the header's resolution line; offsets_host[H + 1];
 * *rle_out = 1 (run-length coded scanlines) or 0 (flat RGBE).  One sequential pass over the packet headers only. */
int uncl_hdr_scan_host(const unsigned char* data_host, long size, long pixel_offset, int W, int H, long* offsets_host,
                       int* rle_out);
/* Expand the scanlines on the device (one CTA per scanline) and convert RGBE to float exactly as Radiance / OpenCV do
 * (mantissa * 2^(e - 136)).  data / offsets: DEVICE copies of the file bytes and of the scan's offsets; out fp32 [3][H][W]. */
int uncl_hdr_decode(const unsigned char* data, const long* offsets, int W, int H, int rle, float* out, uncl_stream_t stream);

/* ---- off-default-path operators ---- */

/* adaptive_lambda.cross_entropy (utils/adaptive_lambda.py:7-21) for a population of L candidate lambdas at once:
 * log10(g*lambda+1)/max -> `bins`-bin density histogram over [0,1] -> cross entropy against `targets`.
 * lambdas: L doubles on the device.  workspace: (L*bins + 4)*4 bytes. */
int uncl_lambda_cross_entropy(const float* gray, long n, const double* lambdas, int L, const float* targets, int bins,
                              float* ce_out, void* workspace, uncl_stream_t stream);
/* models/Blocks.py:77-138 on a dense [N][per] tensor.  mode: 0 Exp, 1 MySig(param), 2 Clip, 3 MaxNormalization,
 * 4 MaxNormalizationEpsilon, 5 BatchMaxNormalization, 6 MinMaxNormalization.  scratch: 2*N floats. */
int uncl_blocks_apply(const float* x, float* out, int N, long per, int mode, float param, float* scratch,
                      uncl_stream_t stream);

/* ---- training-sample preparation on the device ----
 * utils/ProcessedDatasetFolderImg.py:44-168, utils/ProcessedDatasetFolder.py:43-215 (npy_loader), :13-22 (get_ldr_im). */

/* cv2.resize (INTER_LINEAR, float32) of an HWC image to RH x RW followed by the P x P crop at (xx, yy), written CHW.
 * src_base / src_off: the batch's source images packed in one device buffer, per-crop offsets in floats;
 * meta: int32 [S][8] = {H, W, RH, RW, xx, yy, 0, 0} (RH == H and RW == W: no resize); color [S][3][P][P]. */
int uncl_sample_crop_resize(const float* src_base, const long* src_off, const int* meta, float* color, int S, int P,
                            uncl_stream_t stream);
/* Y = .299R + .587G + .114B and the loader's normalisation.  mode 0 (HDR): input = log10((Y-min)/max(Y-min)*f+1) /
 * max(...), gray_norm = Y/max, gray_shift = Y-min (either may be NULL); mode 1: Y/max; mode 2: Y/255;
 * mode 3: clip(((Y-min)/max)*max_stretch - min_stretch, 0, 1).  f_per_sample: S floats (mode 0).  stats_scratch: 2*S. */
int uncl_sample_normalise(const float* color, int S, int P, int mode, const float* f_per_sample, float max_stretch,
                          float min_stretch, float* input, float* gray_norm, float* gray_shift, float* stats_scratch,
                          uncl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* UNCLTMO_B200_H */
