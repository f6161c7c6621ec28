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
 * w_packed: bf16 [C_in/16][9][2][C_out][8] (see uncltmo_b200/packing.py); bias fp32.
 * fuse_outc=1: apply the 1x1 out conv (outc_w [C_out], outc_b [1]) + sigmoid in the epilogue and write
 * out_img [N][Ho][Wo] fp32 (unet_parts.py:338-345 + nn.Sigmoid, Unet_singleFrame.py:207-209); `out` may then be
 * NULL to skip storing the feature map. */
int uncl_conv3x3_tc(const void* in, long in_img_stride, const void* w_packed, const float* bias, void* out,
                    long out_img_stride, int N, int C_in, int H, int W, int C_out, int pad, int act, int emit_skip,
                    int fuse_outc, const float* outc_w, const float* outc_b, float* out_img, float* out_logit,
                    uncl_stream_t stream);

/* up.up: nn.ConvTranspose2d(C, C, 2, stride=2) + bias, written into an (H2 x W2) channel slice of the skip
 * concat buffer with F.pad(..., mode='replicate') semantics.  unet_parts.py:283-299.  w [C][4][C] fp32.
 * prev/r: video generator recurrence - the first r input channels come from `prev` (Unet.py:270). */
int uncl_convT2x2(const void* in, long in_img_stride, const void* prev, long prev_img_stride, int r, const float* w,
                  const float* bias, void* out, long out_img_stride, int N, int C, int H, int W, int H2, int W2,
                  int dtype, uncl_stream_t stream);

/* down.mpconv[0]: nn.MaxPool2d(2).  unet_parts.py:210-213.  prev/r as above (Unet.py:244). */
int uncl_maxpool2(const void* in, long in_img_stride, const void* prev, long prev_img_stride, int r, void* out,
                  long out_img_stride, int N, int C, int H, int W, int dtype, uncl_stream_t stream);

/* outconv 1x1 (C -> 1) + nn.Sigmoid.  unet_parts.py:338-345, Unet_singleFrame.py:207-209.
 * out / logit: [N][HW] fp32 (logit may be NULL). */
int uncl_outc_sigmoid(const void* in, long in_img_stride, const float* w, const float* b, float* out, float* logit,
                      int N, int C, int HW, int dtype, uncl_stream_t stream);

/* layout transforms at the module boundary (NCHW fp32 <-> blocked) */
int uncl_blocked_to_nchw(const void* in, long in_img_stride, float* out, int N, int C, int HW, int dtype,
                         uncl_stream_t stream);
int uncl_nchw_to_blocked(const float* in, void* out, long out_img_stride, int N, int C, int HW, int dtype,
                         uncl_stream_t stream);

/* ---- bottleneck graph block (Unet_singleFrame.py:44-99, gcn_lib/) - fp32 blocked tensors, HW = 144 ---- */

/* x + pos_embed.  Unet_singleFrame.py:94.  pos: blocked fp32 [C/8][144][8]. */
int uncl_gcn_add_pos(const void* in, long in_img_stride, const float* pos, float* out, int N, int C, int dtype,
                     uncl_stream_t stream);

/* 1x1 conv (+groups, bias, act, residual, per-sample DropPath scale): Grapher fc1/fc2, BasicConv (groups=4), FFN.
 * gcn_lib/torch_vertex.py:219-227, gcn_lib/torch_nn.py:54-78, Unet_singleFrame.py:36-42.
 * w [groups][C_in/groups][C_out/groups] fp32; out = scale[n] * act(conv + bias) + res. */
int uncl_pw_conv(const float* in, const float* w, const float* bias, const float* res, const float* scale, void* out,
                 long out_img_stride, int N, int C_in, int C_out, int groups, int HW, int act, int out_dtype,
                 uncl_stream_t stream);

/* DenseDilatedKnnGraph (k=9, dilation 1, relative_pos) + MRConv2d aggregation + channel interleave.
 * gcn_lib/torch_edge.py:135-159, 54-86, 9-20; gcn_lib/torch_vertex.py:21-30.
 * y [N][C/8][144][8]; relpos [144][144]; z [N][2C/8][144][8]; idx_out [N][144][9] int32 or NULL. */
int uncl_gcn_knn_aggregate(const float* y, const float* relpos, float* z, int* idx_out, int N, int C,
                           uncl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* UNCLTMO_B200_H */
