/*
 * dfir.h — C ABI of libdfir_b200.so: the B200 (sm_100a) implementation of the Deep-FIR SISR forward hot path.
 *
 * The upstream reference (um-dsrg/Super-Resolution-Meta-Attention-Networks) is pure Python/PyTorch and has no
 * FFI of its own; its boundary for this path is `net.forward(x, metadata)` of the nn.Modules that the
 * Q-model handlers own.  Each entry point below replaces one reference function and cites
 * it (paths relative to /root/reference/Code/SISR/models).  INTEGRATION.md shows the ctypes binding a
 * reference maintainer would add.
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / C++ types.  All data pointers are DEVICE pointers unless a name
 *     ends in `_host`.  `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - every function returns DFIR_OK (0) or a negative DFIR_ERR_* code and never throws; work is enqueued
 *     on `stream` and is asynchronous; nothing allocates device memory (the caller passes workspaces).
 *   - no global mutable state: all state lives in the caller-owned descriptor structs and buffers.
 *   - activations inside the library: NHWC; images at the API edge: NCHW fp32 contiguous, like the
 *     reference's tensors.
 */
#ifndef DFIR_H_
#define DFIR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFIR_ABI_VERSION 2

/* error codes */
#define DFIR_OK 0
#define DFIR_ERR_ARG (-1)       /* invalid argument / unsupported configuration */
#define DFIR_ERR_CUDA (-2)      /* a CUDA runtime call or kernel launch failed  */
#define DFIR_ERR_DRIVER (-3)    /* cuTensorMapEncodeTiled entry point missing   */
#define DFIR_ERR_TMAP (-4)      /* tensor-map encoding rejected                 */
#define DFIR_ERR_WORKSPACE (-5) /* workspace too small                          */
#define DFIR_ERR_ARCH (-6)      /* device is not sm_100                         */

/* channel-attention styles of QCALayer (attention_manipulators/architectures.py:34-127) */
#define DFIR_STYLE_NONE 0 /* no channel attention (ParamResBlock / QRB) */
#define DFIR_STYLE_STANDARD 1
#define DFIR_STYLE_MODULATE 2
#define DFIR_STYLE_MAX_CONCAT 3
#define DFIR_STYLE_SOFTMAX 4
#define DFIR_STYLE_MINI_CONCAT 5
#define DFIR_STYLE_EXTENDED 6

/* arithmetic modes */
#define DFIR_PREC_BF16_TC 0  /* bf16 operands, fp32 accumulate on tcgen05; fp32 residual stream */
#define DFIR_PREC_FP32_SIMT 1 /* fp32 everywhere on CUDA cores (parity mode: <= 1e-4 normalised error) */

const char* dfir_version(void);
const char* dfir_error_string(int code);
/* 0 if the current device can run the library (compute capability 10.x), else DFIR_ERR_ARCH */
int dfir_check_device(void);
/* Synchronises the device and copies the 8-word pipeline watchdog record to out8_host ({0,...} = no barrier wait
 * ever timed out; else {1, wait tag, blockIdx, threadIdx, parity, barrier lo, barrier hi, 0}); optionally clears it. */
int dfir_debug_watchdog(unsigned int* out8_host, int reset);
/* Synchronises the device and copies the pipeline time-stamp record of the conv kernel ([16 event kinds][64 rows] SM clock
 * values of one CTA) to out1024_host.  Only a `make PROBES=1` build with DFIR_DEBUG_PROBE bit 32768 fills it. */
int dfir_debug_trace(unsigned long long* out1024_host);

/* ------------------------------------------------------------------------------------------------
 * weight packing (derived caches of the fp32 OIHW nn.Parameters; never serialised)
 * ---------------------------------------------------------------------------------------------- */

/* OIHW fp32 [cout][cin=64][3][3] -> tensor-core layout: [9 taps][nt_rows][64 cin] bf16, rows in the UMMA
 * K-major SWIZZLE_128B byte order.  Row n holds output channel co_begin + n*co_stride (zero rows beyond
 * cout).  nt_rows = 64 (trunk/upsampler slices) or 16 (tail).  out: 9*nt_rows*128 bytes. */
int dfir_pack_conv3x3_bf16(const float* w_oihw, void* out, int cout, int cin, int nt_rows, int co_begin,
                           int co_stride, void* stream);
/* OIHW fp32 -> [9 taps][cin][cout] fp32 for the CUDA-core kernels. */
int dfir_pack_conv3x3_f32(const float* w_oihw, float* out, int cout, int cin, void* stream);

/* ------------------------------------------------------------------------------------------------
 * single operators (each is also exercised on its own by tests/test_ops_gpu.py)
 * ---------------------------------------------------------------------------------------------- */

/* default_conv (advanced/common.py:5-8), 64 -> 64 (or 64 -> <=16 with epi = 4) on the tensor cores.
 *   in_bf16  : NHWC bf16 [B][H][W][cin_total]; the 64 channels starting at cin_off are consumed
 *   wpacked  : from dfir_pack_conv3x3_bf16;  bias: fp32 [64] (or [16])
 *   epi      : 0 bias | 1 bias+ReLU | 2 bias + per-row channel sums into pool_rows[B][nseg][H][64]
 *              | 3 bias + skip_f32 (NHWC fp32) -> out_f32 (NHWC fp32, may be NULL) and out_bf16 (TMA store;
 *                works with strided outputs; dfir_conv3x3_c64_scale_skip is the fast dense variant)
 *              | 4 tail: out_f32 is NCHW fp32 [B][cout][H][W], no bf16 output
 *   out_bf16 : NHWC bf16, addressed with explicit byte strides so that PixelShuffle
 *              (advanced/common.py:30) folds into the store: for sub-pixel (i,j) of an r-times upsampler
 *              pass base + ((i*r*W + j)*64*2), pix stride r*128, row stride r*(r*W)*128.
 *   desc_mode: reserved, must be 0 (an alternative shared-memory descriptor encoding used during hardware
 *              bring-up lived here; see DESIGN.md "descriptor base_offset"). */
int dfir_conv3x3_c64(const void* in_bf16, int cin_total, int cin_off, const void* wpacked, const float* bias,
                     int B, int H, int W, int epi, int cout, void* out_bf16, long long out_pix_stride,
                     long long out_row_stride, long long out_img_stride, const float* skip_f32, float* out_f32,
                     float* pool_rows, int desc_mode, void* stream);

/* Conv whose operand is formed on the fly from the previous block (the channel-attention / meta-attention
 * scale and the residual add of QRCAB, attention_manipulators/architectures.py:105-127,172-180; q_layer.py:43;
 * ParamResBlock :346-356 with style NONE):
 *     s[b]  = CA_style(mean over pixels of r_b, attributes[b]) * (sq ? sq[b] : 1)   (style NONE: res_scale * sq)
 *     x'    = r * s[b] + x_in ;   out = default_conv(bf16(x'))
 *   r: NHWC bf16; pool_rows[B][nseg][H][64]: per-row channel sums of r (from dfir_conv3x3_c64 epi 2);
 *   x_in: NHWC fp32 residual stream; x_out (optional, must not alias x_in): NHWC fp32 copy of x';
 *   ca_params: the block's QCALayer parameters (DESIGN.md "attention parameter blob"), R = 64/reduction.
 *   epi: 1 (bias+ReLU -> out_bf16) or 3 (bias + skip_f32 -> out_f32 [optional] and out_bf16). */
int dfir_conv3x3_c64_fused(const void* r_bf16, const float* x_in, float* x_out, const float* pool_rows, int style,
                           const float* ca_params, int R, int M, int A, const float* attributes, const float* sq,
                           float res_scale, const void* wpacked, const float* bias, int B, int H, int W, int epi,
                           void* out_bf16, const float* skip_f32, float* out_f32, void* stream);

/* --- pool-by-linearity trio (the default Q-RCAN schedule) -------------------------------------------------
 * The channel attention of an RCAB needs mean_pixels(r), r = conv2(t).  That mean is linear in t, so it
 * follows from a few sums of t = relu(conv1(x)) — before conv2 runs — and conv2's epilogue can apply the
 * attention scale and the residual add itself (QRCAB.forward, attention_manipulators/architectures.py:172-180).
 *
 * dfir_conv3x3_c64_stats : RCAB conv1 (`body[0]` + ReLU, :155-160): out_bf16 = t (dense NHWC) and, of the
 *   bf16-rounded t: pool_rows[B][nseg][H][64] (per-row channel sums), col_first/col_last[B][H][64] (t at x = 0
 *   and x = W-1 of every row).
 * dfir_ca_from_stats     : QCALayer (:105-125) on mean(conv2(t)) reconstructed from those sums, times the
 *   meta-attention scale sq (q_layer.py:39-43): svec[B][64].  w2_packed/bias2 = conv2's packed weights / bias.
 * dfir_conv3x3_c64_scale_skip : RCAB conv2 + `res * y` (twice) + `res += x` (:173-179), also the group / trunk
 *   tail conv (:231-232, :312-313): out_f32 = (conv(in) + bias) * s[b] + skip_f32 (NHWC fp32, may alias skip_f32,
 *   may be NULL), out_bf16 = bf16(out_f32) (dense NHWC).  The scale s is, in this order of precedence:
 *   style != DFIR_STYLE_NONE -> evaluated INSIDE the kernel from pool_rows/col_first/col_last (what
 *   dfir_ca_from_stats computes, times sq) while the pipeline fills; else svec[B][64] if not NULL; else 1. */
int dfir_conv3x3_c64_stats(const void* in_bf16, const void* wpacked, const float* bias, int B, int H, int W,
                           void* out_bf16, float* pool_rows, float* col_first, float* col_last, void* stream);
/* dfir_conv3x3_c64_stats with the statistics as the 64-bit FIXED-POINT image sums dfir_qrcan_forward uses (2^-24 units,
 * accumulated with integer atomics, so independent of how rows are grouped: batch-invariant).  istats [B][9][64] int64,
 * zeroed by the caller, entries: total, first / last column sums, first / last row sums, the four corner pixels (0,0), (0,W-1),
 * (H-1,0), (H-1,W-1).  warp_autonomous != 0: the epilogue warps stage and store 32 pixels each and keep 32-bit per-thread
 * sums (2^-12 units) that are reduced when the image changes (not for W with (W - 1) % 128 in {0, 8, 16, 24}). */
int dfir_conv3x3_c64_stats_fx(const void* in_bf16, const void* wpacked, const float* bias, int B, int H, int W,
                              void* out_bf16, long long* istats, int warp_autonomous, void* stream);
int dfir_ca_from_stats(const float* pool_rows, const float* col_first, const float* col_last, const void* w2_packed,
                       const float* bias2, int style, const float* ca_params, int R, int M, int A,
                       const float* attributes, const float* sq, float* svec, int B, int H, int W, void* stream);
int dfir_conv3x3_c64_scale_skip(const void* in_bf16, const void* wpacked, const float* bias, int B, int H, int W,
                                const float* svec, const float* skip_f32, float* out_f32, void* out_bf16,
                                const float* pool_rows, const float* col_first, const float* col_last, int style,
                                const float* ca_params, int R, int M, int A, const float* attributes,
                                const float* sq, void* stream);
/* The same op on the bf16 hi + bf16 lo residual stream of the inference chain (x = hi + lo, 16 significant bits; the hi
 * plane is the operand of the next conv): out = (conv(in) + b) * s + (skip_hi + skip_lo), out_hi = bf16(out), out_lo =
 * bf16(out - out_hi) (out_lo may be NULL when the result only feeds a conv).  All planes dense NHWC bf16; out may alias
 * skip (in place).  10 instead of 12 bytes of HBM traffic per element; tiles travel by TMA.  Reference: the same lines as
 * dfir_conv3x3_c64_scale_skip (architectures.py:105-127,172-180; q_layer.py:39-43).
 * descending != 0 (W <= 128 only): images and rows are traversed from the last to the first — same result up to the fp32
 * summation order of the taps; use it after an ascending producer of `in` so that its last rows are read from L2. */
int dfir_conv3x3_c64_scale_skip_hl(const void* in_bf16, const void* wpacked, const float* bias, int B, int H, int W,
                                   const float* svec, const void* skip_hi, const void* skip_lo, void* out_hi, void* out_lo,
                                   const float* pool_rows, const float* col_first, const float* col_last, int style,
                                   const float* ca_params, int R, int M, int A, const float* attributes,
                                   const float* sq, int descending, void* stream);
/* The same with an 8-BIT lo plane (the format dfir_qrcan_forward uses by default).  The stream value is a 24-bit float X
 * (sign, 8 exponent bits, 15 mantissa bits: 16 significant bits, like hi + lo above) stored as two planes whose BIT PATTERNS
 * add up: bits(X) = (hi << 16) + (q << 8) with hi = the bf16 nearest to X (ties away from zero) and q = int8 in [-128, 127].
 * X = the result rounded to 24 bits (ties away from zero).  The integer form is exact across binade boundaries and needs no
 * exponent arithmetic in the epilogue.  8 bytes of HBM traffic per element (t 2, hi 2 + 2, lo 1 + 1).
 * skip_lo8 / out_lo8: dense int8 planes [B][H][W][64] whose 64 bytes per pixel are stored in the order of the accumulator
 * fragment: byte cq * 16 + 2 n + e holds channel 8 n + 2 cq + e (n = 0..7, cq = 0..3, e = 0..1), so that a thread of the
 * epilogue reads its 16 channels of a pixel as one 16-byte word.  The plane is private to the library (dfir_stream_encode_hl8 /
 * dfir_stream_decode_hl8 convert from / to fp32).
 * Statistics for the in-kernel attention (style != NONE), two forms: the three per-row arrays of dfir_conv3x3_c64_stats
 * (pool_rows, col_first, col_last), or - what dfir_qrcan_forward launches - the [B][9][64] int64 fixed-point image sums of
 * dfir_conv3x3_c64_stats_fx passed through `pool_rows` with col_first == col_last == NULL (both hl entry points). */
int dfir_conv3x3_c64_scale_skip_hl8(const void* in_bf16, const void* wpacked, const float* bias, int B, int H, int W,
                                   const float* svec, const void* skip_hi, const void* skip_lo8, void* out_hi, void* out_lo8,
                                   const float* pool_rows, const float* col_first, const float* col_last, int style,
                                   const float* ca_params, int R, int M, int A, const float* attributes,
                                   const float* sq, int descending, void* stream);

/* fp32 <-> the hi / 8-bit lo stream format above: hi [n] bf16, lo8 [n] int8, x [n] fp32 (n a multiple of 4).  decode(encode(x))
 * = x rounded to 24 bits (ties away from zero). */
int dfir_stream_encode_hl8(const float* x, void* hi, void* lo8, long long n, void* stream);
int dfir_stream_decode_hl8(const void* hi, const void* lo8, float* x, long long n, void* stream);

/* K-chunked convolutions for feature widths above 64 (Q-EDSR with 128/192/256 features, architectures.py:359-399):
 * a C -> C conv is a (C/64) x (C/64) block matrix of 64 -> 64 convs over 64-channel planes; the input-chunk sums are
 * accumulated in fp32 by re-launching the tensor-core kernel with the running sum as its skip input:
 *     out_f32 = conv(in) * svec[b] + bias * svec[b] + skip_f32  [ReLU]      (svec NULL = 1, bias/skip NULL = 0)
 *     out_bf16 = bf16(out_f32)       (dense NHWC, 64 channels; out_f32 may be NULL and may alias skip_f32)
 * With svec = res_scale * meta scale and skip = the fp32 stream this is ParamResBlock's `res * y + x` (:352-355)
 * accumulated in place. */
/* dfir_conv3x3_c64_accumulate with the running sum kept in the hi / 8-bit lo stream format (16 significant bits per
 * accumulation step instead of fp32, 8 instead of 14 bytes of HBM traffic per element and the TMA-tile epilogue of
 * dfir_conv3x3_c64_scale_skip_hl8):   out = conv(in) * svec[b] + bias * svec[b] + (skip_hi, skip_lo8)  [ReLU];
 * out_hi = the bf16 operand of the next conv, out_lo8 may be NULL when only that is needed; out may alias skip. */
int dfir_conv3x3_c64_accumulate_hl8(const void* in_bf16, const void* wpacked, const float* bias, int B, int H, int W,
                                    const float* svec, const void* skip_hi, const void* skip_lo8, void* out_hi, void* out_lo8,
                                    int relu, void* stream);
int dfir_conv3x3_c64_accumulate(const void* in_bf16, const void* wpacked, const float* bias, int B, int H, int W,
                                const float* svec, const float* skip_f32, float* out_f32, void* out_bf16, int relu,
                                void* stream);
/* tail conv 64 -> cout (<= 16) writing / accumulating into fp32 NCHW [B][cout][H][W]; wpacked from
 * dfir_pack_conv3x3_bf16 with nt_rows = 16 */
int dfir_conv3x3_c64_tail(const void* in_bf16, const void* wpacked, const float* bias, int B, int H, int W, int cout,
                          float* out_nchw, int accumulate, void* stream);

/* default_conv on CUDA cores, fp32 NHWC in/out, any Cin % 4 == 0 and any Cout.
 *   w_packed from dfir_pack_conv3x3_f32; skip (optional) NHWC fp32 added after bias; relu applied last;
 *   ps_r > 1 folds nn.PixelShuffle(ps_r) into the store (out is then [B][H*r][W*r][Cout/r^2]);
 *   out_nchw != 0 writes [B][Cout][H][W] instead. */
int dfir_conv3x3_f32(const float* in, const float* w_packed, const float* bias, const float* skip, float* out, int B,
                     int H, int W, int Cin, int Cout, int relu, int ps_r, int out_nchw, void* stream);

/* head conv (QRCAN.head, attention_manipulators/architectures.py:275,310): NCHW fp32 image -> NHWC features,
 * fp32 (out_f32) and/or bf16 (out_bf16) copies; w_packed from dfir_pack_conv3x3_f32. */
int dfir_head_conv(const float* x_nchw, const float* w_packed, const float* bias, float* out_f32, void* out_bf16,
                   int B, int Cin, int H, int W, int Cout, void* stream);

/* ParaCALayer.attribute_integrator for nblk layers at once (attention_manipulators/q_layer.py:21-41):
 *   out[blk][b][:] = sigmoid(W2[blk] * act(W1[blk] * meta[b] + b1[blk]) + b2[blk]), act = ReLU if relu else id.
 *   meta fp32 [B][M]; w1 [nblk][Hid][M]; b1 [nblk][Hid]; w2 [nblk][C][Hid]; b2 [nblk][C]; out [nblk][B][C].
 *   blk_enabled (optional, int32 [nblk]): 0 -> the layer is absent and out = 1.  out_scale multiplies every
 *   output (ParamResBlock's res_scale folded into the meta scale; 1 otherwise). */
int dfir_meta_attention(const float* meta, const float* w1, const float* b1, const float* w2, const float* b2,
                        float* out, int nblk, int B, int M, int Hid, int C, int relu, const int* blk_enabled,
                        float out_scale, void* stream);

/* Channel attention + meta-attention scale + residual add of one block:
 *   QCALayer.forward + ParaCALayer `x*y` + `res += x` (attention_manipulators/architectures.py:105-127,172-180;
 *   q_layer.py:43; ParamResBlock :346-356 with style NONE and res_scale):
 *     y  = mean over pixels of r, rebuilt from pool_rows[b][0..pool_nrows)[C] / (H*W)
 *     s  = CA_style(y, attributes) (* sq[b][:] if sq != NULL)            (style NONE: s = res_scale * sq)
 *     x_out = r * s + x_in     (fp32 NHWC; x_out may alias x_in), x_out_bf16 = bf16(x_out) (optional)
 *   r is NHWC bf16 (r_is_bf16) or fp32; ca_params: the block's fp32 parameter arrays, concatenated in the
 *   order documented in DESIGN.md §"attention parameter blob". */
int dfir_ca_scale_residual(const void* r, int r_is_bf16, const float* x_in, const float* pool_rows, int pool_nrows,
                           int style, const float* ca_params, int C, int R, int M, int A, const float* attributes,
                           const float* sq, float res_scale, float* x_out, void* x_out_bf16, int B, int H, int W,
                           void* stream);
/* The same with PALayer between the channel attention and the meta attention (QRCAB.forward,
 * attention_manipulators/architectures.py:172-180 with include_pixel_attention):
 *     u = r * CA(y) ;  u <- u * sigmoid(W2 relu(W1 u + b1) + b2)  (per pixel) ;  x_out = u * sq + x_in
 * pa_params: W1[8][64] b1[8] W2[8] b2[1] (fp32); C must be 64, style != none. */
int dfir_ca_pa_scale_residual(const void* r, int r_is_bf16, const float* x_in, const float* pool_rows, int pool_nrows,
                              int style, const float* ca_params, const float* pa_params, int R, int M, int A,
                              const float* attributes, const float* sq, float* x_out, void* x_out_bf16, int B, int H,
                              int W, void* stream);

/* Post-processing of an SR batch on the device — what ModelInterface.net_run_and_process does with numpy on the host
 * (models/__init__.py:138-169; sr_tools/image_manipulation.py:56-157, im_type 'jpg'):
 *   rgb_clipped = clip(x, lo, hi);  ycbcr = [0.299 r + 0.587 g + 0.114 b,  128/255 + (-0.168736 r - 0.331264 g + 0.5 b),
 *                                            128/255 + (0.5 r - 0.418688 g - 0.081312 b)]  of the clipped values.
 * x, rgb_clipped, ycbcr: fp32 NCHW [B][3][HW].  Bit-identical to the numpy expressions (separately rounded fp32 ops). */
int dfir_postprocess_rgb(const float* x_nchw, float* rgb_clipped, float* ycbcr, int B, long long HW, float lo, float hi,
                         void* stream);

/* The same post-processing with 8-bit results and the Y-channel PSNR on the device: rgb_u8 / ycbcr_u8
 * = rint(255 * clip(., 0, 1)) of the arrays dfir_postprocess_rgb produces (either may be NULL); when hr_nchw (the ground
 * truth, fp32 NCHW in [0,1]) is given, y_psnr[b] = 20 log10(1 / sqrt(mean((Y_sr - Y_hr)^2))) of image b, 100 for identical
 * images, exactly what Metrics.run_image_metric('PSNR') computes from channel 0 of the YCbCr arrays (sr_tools/metrics.py:6-17,
 * max_value 1).  6 B written per pixel instead of 24; scratch: dfir_postprocess_u8_scratch_bytes (only with hr_nchw). */
size_t dfir_postprocess_u8_scratch_bytes(int B, long long HW);
int dfir_postprocess_u8(const float* x_nchw, const float* hr_nchw, unsigned char* rgb_u8, unsigned char* ycbcr_u8,
                        float* y_psnr, void* scratch, size_t scratch_bytes, int B, long long HW, void* stream);

/* per-row channel sums of an fp32 NHWC tensor: pool_rows[b][y][c] = sum_x in[b][y][x][c] (fp32 mode only;
 * the tensor-core conv produces them in its epilogue). */
int dfir_pool_rows_f32(const float* in, float* pool_rows, int B, int H, int W, int C, void* stream);

/* ------------------------------------------------------------------------------------------------
 * whole-network forward
 * ---------------------------------------------------------------------------------------------- */

/* Packed parameters + structure of a Q-RCAN (attention_manipulators/architectures.py:246-316) or, with
 * style = DFIR_STYLE_NONE, n_groups = 1 and no_group_conv = 1, of a Q-EDSR (:359-399: ParamResBlock chain).
 * Conv indexing of the trunk arrays (P = 2*n_blocks + 1, or 2*n_blocks when no_group_conv): block (g,b) conv j
 * -> g*P + 2*b + j; group tail conv -> g*P + 2*n_blocks; trunk tail conv (final_body) -> n_groups*P;
 * then the upsampler slices: stage t, sub-pixel s -> n_trunk + t*r*r + s.                                  */
typedef struct dfir_qrcan_net {
  int n_groups, n_blocks, n_feats; /* n_feats must be 64 */
  int scale;                        /* 2, 3, 4, 8 */
  int style;                        /* DFIR_STYLE_* */
  int reduced;                      /* n_feats / reduction */
  int num_metadata;                 /* M: size of the metadata vector the FC layers were built for */
  int attr_size;                    /* A: size of the `attributes` rows passed at run time (64 for modulate) */
  int meta_hidden;                  /* hidden width of ParaCALayer (n_feats/2 for M <= 15) */
  int in_feats, out_feats;          /* 3, 3 */
  const int* q_enabled;             /* device int32 [n_groups*n_blocks]: block owns a q_node */
  int any_q;                        /* host-side: any block has a q_node */
  int chunk_images;                 /* images per L2-resident pass; 0 = choose automatically */
  int schedule;                     /* block chain: 0 pool-by-linearity (default), 1 fused-in, 2 streamer, 3 pool-by-linearity with
                                       the attention vector from its own kernel */
  int no_group_conv;                /* 1: groups have no tail conv / group skip (Q-EDSR: one flat chain of blocks) */
  int meta_relu;                    /* ReLU between the two FC layers of the meta-attention MLP (q_layer.py:33-34) */
  float res_scale;                  /* ParamResBlock res_scale (style NONE only; QRCAB ignores it) */
  /* tensor-core weights */
  const void* conv_w_bf16;          /* [n_conv][9*64*128 B] */
  const void* tail_w_bf16;          /* [9*16*128 B] */
  /* CUDA-core weights (fp32 mode; may be NULL when only bf16 mode is used) */
  const float* conv_w_f32;          /* trunk: [n_trunk][9][64][64]; then the unsliced upsampler convs */
  const float* up_w_f32;            /* [n_up][9][64][r*r*64] */
  const float* tail_w_f32;          /* [9][64][out_feats] */
  const float* head_w_f32;          /* [9][in_feats][64] */
  /* biases, fp32 */
  const float* conv_b;              /* [n_conv][64] (upsampler slices permuted like the weights) */
  const float* up_b;                /* [n_up][r*r*64] in original channel order (fp32 mode) */
  const float* tail_b;              /* [16] (zero padded) */
  const float* head_b;              /* [64] */
  /* attention parameters */
  const float* ca_blob;             /* [n_groups*n_blocks][ca_stride] */
  int ca_stride;
  const float* meta_w1; const float* meta_b1; const float* meta_w2; const float* meta_b2;
  /* training only (may be NULL for inference): weights of the data-gradient convolutions, i.e. the same layouts
   * built from the transposed, 180-degree rotated kernels */
  const void* conv_wT_bf16;         /* [n_conv][9*64*128 B]: trunk convs, then the upsampler slices */
  const float* conv_wT_f32;         /* [n_trunk][9][C][C] */
  const float* up_wT_f32;           /* [n_up][9][r*r*C][C] */
  const float* tail_wT_f32;         /* [9][out_feats][C] (both precisions: the 3 -> C gradient conv runs on CUDA cores) */
  /* PALayer (attention_manipulators/architectures.py:13-26), present when include_pixel_attention: per block
   * pa.0.weight[8][64], pa.0.bias[8], pa.2.weight[8], pa.2.bias[1]; NULL = no pixel attention.  n_feats must be 64;
   * the block chain then runs the streamer schedule (schedule is ignored). */
  const float* pa_blob;             /* [n_groups*n_blocks][pa_stride] */
  int pa_stride;
} dfir_qrcan_net;

size_t dfir_qrcan_workspace_bytes(const dfir_qrcan_net* net, int B, int H, int W, int precision);
/* number of kernel launches one dfir_qrcan_forward call enqueues (bench.py's `gpu_launches` claim) */
long long dfir_qrcan_launch_count(const dfir_qrcan_net* net, int B, int H, int W, int precision);

/* QRCAN.forward(x, metadata) (attention_manipulators/architectures.py:309-316).
 *   x_nchw fp32 [B][3][H][W]; attributes fp32 [B][attr_size] (the (B,M,1,1) tensor of QModel.generate_channels);
 *   out_nchw fp32 [B][3][scale*H][scale*W].  workspace >= dfir_qrcan_workspace_bytes(). */
int dfir_qrcan_forward(const dfir_qrcan_net* net, const float* x_nchw, const float* attributes, float* out_nchw,
                       int B, int H, int W, int precision, void* workspace, size_t workspace_bytes, void* stream);

/* Staged execution of the same schedule (the state lives in the caller's workspace; requires one pass, i.e.
 * chunk_images = 0/B and a workspace for B images).  stages is a mask: 1 head conv, 2 groups [g_begin, g_end),
 * 4 trunk tail conv + head skip, 8 upsampler + tail conv.  Used by the host code of Q-HAN / Q-SAN, whose extra
 * layers sit between these stages (attention_manipulators/architectures.py:447-467, 514-540).
 *   feat_in_f32  (optional, NHWC fp32): replaces the stream entering g_begin (stages & 2), the trunk-tail input
 *                (stages & 4 without 2) or the upsampler input (stages == 8)
 *   group_out_f32(optional): fp32 NHWC copy of the stream after every executed group, [g_end-g_begin][B][H][W][C]
 *   feat_out_f32 (optional): fp32 NHWC copy of the stream after the last executed group, or of the head output */
int dfir_qrcan_stages(const dfir_qrcan_net* net, int stages, int g_begin, int g_end, const float* x_nchw,
                      const float* attributes, const float* feat_in_f32, float* group_out_f32, float* feat_out_f32,
                      float* out_nchw, int B, int H, int W, int precision, void* workspace, size_t workspace_bytes,
                      void* stream);

/* ------------------------------------------------------------------------------------------------
 * training step: forward with saved activations + backward
 *   (BaseModel.run_train / standard_update, models/__init__.py:466-489: forward, L1 loss, loss.backward(), Adam)
 * The loss, the optimizer and the scheduler stay in PyTorch (they own the fp32 nn.Parameters); the library computes
 * the network forward and every parameter gradient.  Supported: Q-RCAN with every channel-attention style (no pixel
 * attention) and Q-EDSR (flat chain), n_feats = 64 on the tensor-core path, any width in fp32 mode.
 * ---------------------------------------------------------------------------------------------- */

/* Device pointers to the fp32 nn.Parameter storages of a network — or, with the same shape, to their gradient
 * buffers.  `conv_w` ... `meta` are arrays of pointers that live in DEVICE memory (the kernels read them), indexed
 * like the trunk arrays of dfir_qrcan_net.  ca holds 8 pointers per block: (weight, bias) of the QCALayer's up to four FC
 * layers in forward order (standard / modulate / max_concat / softmax: conv_du.0, conv_du.2; mini_concat: pre_concat,
 * conv_du.1; extended_attention: feature_convs.{0,1,2}.0, final_conv.0; unused entries NULL).  meta holds 4 pointers per
 * block: ParaCALayer.attribute_integrator {FC1 weight, bias, FC2 weight, bias} (NULL entries for blocks without a q
 * layer).  All tensors keep the reference's OIHW / [out][in] layouts. */
typedef struct dfir_qrcan_params {
  float* const* conv_w;  /* [n_trunk] OIHW [C][C][3][3] */
  float* const* conv_b;  /* [n_trunk] [C] */
  float* const* up_w;    /* [n_up]    OIHW [r*r*C][C][3][3] */
  float* const* up_b;    /* [n_up]    [r*r*C] */
  float* tail_w; float* tail_b; /* [out_feats][C][3][3], [out_feats] */
  float* head_w; float* head_b; /* [C][in_feats][3][3], [C] */
  float* const* ca;      /* [n_groups*n_blocks*8] or NULL (style none) */
  float* const* meta;    /* [n_groups*n_blocks*4] or NULL (no q layers) */
  float* const* pa;      /* [n_groups*n_blocks*4] PALayer.pa {conv 0 weight [8][C], bias [8], conv 2 weight [8], bias [1]}
                            (architectures.py:13-26), or NULL (no pixel attention) */
} dfir_qrcan_params;

/* Rebuilds every kernel-format buffer of `net` (the pointers inside dfir_qrcan_net, written although declared const)
 * from the fp32 parameters in a handful of launches: what must happen after each optimizer.step().
 * with_backward != 0 also fills the conv_wT_* / up_wT_* / tail_wT_* buffers. */
int dfir_qrcan_repack(const dfir_qrcan_net* net, const dfir_qrcan_params* params, int precision, int with_backward,
                      void* stream);

/* Workspace of one training step (activation stash of the forward + scratch of the backward). */
size_t dfir_qrcan_train_workspace_bytes(const dfir_qrcan_net* net, int B, int H, int W, int precision);

/* QRCAN.forward in training mode: same result as dfir_qrcan_forward up to the stated tolerance (bf16 mode: r is
 * rounded to bf16 before the attention scale), every activation the backward needs stays in `workspace`. */
int dfir_qrcan_train_forward(const dfir_qrcan_net* net, const float* x_nchw, const float* attributes, float* out_nchw,
                             int B, int H, int W, int precision, void* workspace, size_t workspace_bytes, void* stream);

/* Backward of the above for grad_out_nchw = dL/d out (fp32 [B][out_feats][sH][sW]).  Writes (not accumulates) the
 * gradient of every parameter through the pointers in `grads`.  Must follow dfir_qrcan_train_forward on the same
 * workspace and the same x / attributes. */
int dfir_qrcan_train_backward(const dfir_qrcan_net* net, const dfir_qrcan_params* grads, const float* x_nchw,
                              const float* attributes, const float* grad_out_nchw, int B, int H, int W, int precision,
                              void* workspace, size_t workspace_bytes, void* stream);
/* kernel launches of one forward + backward (bench.py's gpu_launches claim for the training step) */
long long dfir_qrcan_train_launch_count(const dfir_qrcan_net* net, int B, int H, int W, int precision);

/* Adam.step() (BaseModel.standard_update, models/__init__.py:481-489; optimizer defined at :291-299) over flat fp32
 * buffers: n parameters (n % 4 == 0, 16-byte aligned), torch.optim.Adam arithmetic (L2 weight decay added to the
 * gradient, bias corrections from `step` >= 1), one pass: 16 B read + 12 B written per parameter. */
int dfir_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, long long step, void* stream);

/* single operators of the backward, exercised one by one by tests/test_train_gpu.py */

/* Weight + bias gradient of a 64 -> 64 3x3 conv on the tensor cores.  dy: bf16 NHWC with explicit byte strides
 * (0 = dense); x: dense bf16 NHWC; dw: fp32 OIHW rows co_begin + n*co_stride (n < 64) of a [*][64][3][3] tensor;
 * db likewise.  scratch >= dfir_conv3x3_wgrad_scratch_bytes(). */
size_t dfir_conv3x3_wgrad_scratch_bytes(int B, int H, int W, int Cin, int Cout, int precision);
int dfir_conv3x3_wgrad_c64(const void* dy_bf16, long long dy_pix_stride, long long dy_row_stride,
                           long long dy_img_stride, const void* x_bf16, int B, int H, int W, float* dw_oihw, float* db,
                           int co_begin, int co_stride, void* scratch, size_t scratch_bytes, void* stream);
/* the same on CUDA cores in fp32 for any Cin % 4 == 0 / Cout (dense NHWC operands) */
int dfir_conv3x3_wgrad_f32(const float* dy, const float* x, int B, int H, int W, int Cin, int Cout, float* dw_oihw,
                           float* db, void* scratch, size_t scratch_bytes, void* stream);
/* Data gradient of a 64 -> 64 conv on the tensor cores: out = conv(dy, wT) [* (mask > 0)] [+ skip].
 *   wT_packed: dfir_pack_conv3x3_bf16_ex(transpose = 1).  With mask_bf16 (saved ReLU output) the result is written as
 *   bf16 only; otherwise out_f32 = conv + skip_f32 (skip optional, may alias out_f32) and out_bf16 = its bf16 copy. */
int dfir_conv3x3_c64_dgrad(const void* dy_bf16, long long dy_pix_stride, long long dy_row_stride,
                           long long dy_img_stride, const void* wT_packed, const void* mask_bf16, const float* skip_f32,
                           float* out_f32, void* out_bf16, int B, int H, int W, void* stream);
/* dfir_pack_conv3x3_bf16 with the data-gradient option: transpose != 0 packs rows = input channels, columns = output
 * channels co_begin + k*co_stride, taps mirrored. */
int dfir_pack_conv3x3_bf16_ex(const float* w_oihw, void* out, int cout, int nt_rows, int co_begin, int co_stride,
                              int transpose, void* stream);
/* weight gradients of the 3-channel convs: tail_mode 0 = head conv (img = network input, feat = dL/d head output,
 * fp32), 1 = tail conv (img = dL/d output, feat = tail input; feat_is_bf16 selects its type). */
size_t dfir_conv3x3_wgrad_small_scratch_bytes(int B, int H, int C);
int dfir_conv3x3_wgrad_small(const float* img_nchw, const void* feat_nhwc, int feat_is_bf16, int B, int H, int W, int C,
                             int C3, int tail_mode, float* dw_oihw, float* db, void* scratch, size_t scratch_bytes,
                             void* stream);

/* ------------------------------------------------------------------------------------------------
 * Q-HAN / Q-SAN layers (fp32 NHWC; HBM- or latency-bound kernels, csrc/san_han.cu)
 * ---------------------------------------------------------------------------------------------- */

/* out = x * svec[b][c] (svec optional) + alpha * add (add optional): SOCA's `y * x` (advanced/SAN_blocks.py:302) and
 * the share-source skips `+ gamma * residual` (attention_manipulators/architectures.py:459). */
int dfir_channel_scale(const float* x, const float* svec, const float* add, float alpha, float* out, int B, int HW,
                       int C, void* stream);

/* LAM_Module.forward (advanced/HAN_blocks.py:24-37): N feature maps (map n at stack + n*map_stride_elems, each
 * [B][HW][C] fp32) -> out [B][HW][N*C] with out[..., n*C+c] = gamma * sum_j softmax_j(max_j E[n] - E[n][j]) X_j + X_n,
 * E = Gram matrix over (pixel, channel). */
size_t dfir_lam_scratch_bytes(int B, int N);
int dfir_lam(const float* stack, long long map_stride_elems, float gamma, float* out, void* scratch, int N, int B,
             int HW, int C, void* stream);

/* CSAM_Module.forward (advanced/HAN_blocks.py:59-76): out = x * (gamma * sigmoid(conv3d_3x3x3(x as a (C,H,W)
 * volume) + bias)) + x; w27 = conv.weight[0][0] flattened (channel, y, x). */
int dfir_csam(const float* x, const float* w27, float bias, float gamma, float* out, int B, int H, int W, int C,
              void* stream);

/* SOCA.forward up to its channel scale (advanced/SAN_blocks.py:261-300; Covpool / Sqrtm of advanced/mpncov.py:12-76):
 * covariance pooling over the (centre-cropped when >= 1000) image without materialising the MxM centring matrix,
 * 5 Newton-Schulz iterations, mean over dim 1, FC-ReLU-FC-sigmoid (mlp_params: W1[R][64] b1[R] W2[64][R] b2[64]). */
size_t dfir_soca_scratch_bytes(int B);
int dfir_soca(const float* x, const float* mlp_params, int R, float* svec, void* scratch, int B, int H, int W, int C,
              void* stream);

/* Online degradation of the training input pipeline (Code/sr_tools/gaussian_utils.py:333-424):
 * dfir_batch_blur = BatchBlur.forward (+ the noise / clamp of SRMDPreprocessing.__call__): out[b][c] = reflect-padded
 * x[b][c] cross-correlated with kernels[b] (kernel_per_image != 0, [B][l][l]) or with the one shared kernel ([l][l]),
 * optionally + noise_sigma[b] * noise[b][c] (noise: standard normal samples, same shape as x) and clamped to [0, 1].
 * x, noise, out: fp32 NCHW, out must not alias x; l <= 33, l/2 < min(H, W).
 * dfir_pca_encode = PCAEncoder.__call__: code[b][0..k) = flatten(kernels[b]) @ pca_matrix ([l*l][k], k <= 32); with
 * noise_sigma the code has k + 1 entries and code[b][k] = 10 * noise_sigma[b] (SRMDPreprocessing's `re_code`). */
int dfir_batch_blur(const float* x_nchw, const float* kernels, int kernel_per_image, const float* noise,
                    const float* noise_sigma, float* out_nchw, int B, int C, int H, int W, int l, int clamp01, void* stream);
int dfir_pca_encode(const float* kernels, const float* pca_matrix, const float* noise_sigma, float* code, int B, int l,
                    int k, void* stream);

/* Covpool as a stand-alone operator with the reference's hand-written backward (advanced/mpncov.py:12-47), C = 64:
 *   forward : cov[b] = X (I/M - 11^T/M^2) X^T  of x [B][H][W][64] (NHWC fp32), cov [B][64][64]; crop1000 != 0 applies SOCA's
 *             centre crop to 1000 along sides >= 1000 (SAN_blocks.py:265-280);
 *   backward: grad_x = (G + G^T) X (I/M - 11^T/M^2) for G = grad_cov (zero outside the crop window).
 * Neither materialises the MxM centring matrix (the reference allocates it: 1 GiB at 128x128). */
size_t dfir_covpool_scratch_bytes(int B);
int dfir_covpool(const float* x, float* cov, void* scratch, size_t scratch_bytes, int B, int H, int W, int C, int crop1000,
                 void* stream);
int dfir_covpool_backward(const float* x, const float* grad_cov, float* grad_x, void* scratch, size_t scratch_bytes, int B,
                          int H, int W, int C, int crop1000, void* stream);

/* Sqrtm (Newton-Schulz matrix square root, advanced/mpncov.py:49-112), C = 64, iters >= 2 (the networks use 5):
 *   forward : out[b] = sqrt(tr) * 0.5 Y (3I - Z Y) after iters - 2 coupled iterations on A = cov / tr;
 *   backward: the reference's closed-form gradient of that iteration (Sqrtm.backward), grad_in [B][64][64].
 * One CTA per matrix, everything in shared memory. */
size_t dfir_sqrtm_scratch_bytes(int B, int iters);
int dfir_sqrtm(const float* cov, float* out, int B, int C, int iters, void* stream);
int dfir_sqrtm_backward(const float* cov, const float* grad_out, float* grad_in, void* scratch, size_t scratch_bytes, int B,
                        int C, int iters, void* stream);

/* Nonlocal_CA.forward (advanced/SAN_blocks.py:314-336) with _NonLocalBlockND._embedded_gaussian (:104-148) on each
 * of the 2x2 regions: theta/phi/g 1x1 convs (w_tpg [24][64], b_tpg [24]: theta rows 0-7, phi 8-15, g 16-23),
 * 2x2 max-pool of phi and g (always on), softmax attention, W 1x1 conv (w_out [64][8]) + x. */
size_t dfir_nonlocal_scratch_bytes(int B, int H, int W);
int dfir_nonlocal(const float* x, const float* w_tpg, const float* b_tpg, const float* w_out, const float* b_out,
                  float* out, void* scratch, int B, int H, int W, int C, void* stream);

/* ---- training of the Q-SAN / Q-HAN specific layers (csrc/san_han_bwd.cu).  The reference trains these layers through
 * autograd over the forward code cited at each forward operator above; the functions below restate the derivatives
 * autograd produces.  All tensors fp32, feature maps NHWC, C = 64 unless a C argument says otherwise. */

/* out[b][c] = sum_p a[b][p][c] * b[b][p][c]  (gradient of a per-image channel scale, SAN_blocks.py:302); total (optional,
 * one float) = sum of out over (b, c), added to its previous value when accumulate_total != 0 (gradient of the scalar
 * `gamma` of `x + gamma * residual`, attention_manipulators/architectures.py:459).  out may be NULL. */
size_t dfir_channel_dot_scratch_bytes(int B, int C);
int dfir_channel_dot(const float* a, const float* b, float* out, float* total, int accumulate_total, void* scratch,
                     size_t scratch_bytes, int B, long long HW, int C, void* stream);

/* Tail of SOCA.forward (SAN_blocks.py:290-300) on the square root S [B][64][64]: v = mean over dim 1, svec =
 * sigmoid(W2 relu(W1 v + b1) + b2), and its backward: grad_S [B][64][64], grad_mlp (same flat layout as mlp_params,
 * overwritten; one CTA per image, contributions summed over the batch in index order). */
int dfir_soca_mlp(const float* S, const float* mlp_params, int R, float* svec, int B, void* stream);
size_t dfir_soca_mlp_backward_scratch_bytes(int B, int R);
int dfir_soca_mlp_backward(const float* S, const float* grad_svec, const float* mlp_params, int R, float* grad_S,
                           float* grad_mlp, void* scratch, size_t scratch_bytes, int B, void* stream);

/* Backward of dfir_lam: fwd_scratch is the scratch buffer of the forward call, unmodified (it holds the attention matrix);
 * grad_out [B][HW][N*C]; grad_stack: map n at grad_stack + n * grad_map_stride_elems; grad_gamma: one float (overwritten). */
size_t dfir_lam_backward_scratch_bytes(int B, int N);
int dfir_lam_backward(const float* stack, long long map_stride_elems, const void* fwd_scratch, float gamma,
                      const float* grad_out, float* grad_stack, long long grad_map_stride_elems, float* grad_gamma,
                      void* scratch, size_t scratch_bytes, int N, int B, int HW, int C, void* stream);

/* Backward of dfir_csam: grad_x [B][H][W][C], grad_w27 [27], grad_bias [1], grad_gamma [1] (all overwritten). */
size_t dfir_csam_backward_scratch_bytes(int B, int H, int W, int C);
int dfir_csam_backward(const float* x, const float* grad_out, const float* w27, float bias, float gamma, float* grad_x,
                       float* grad_w27, float* grad_bias, float* grad_gamma, void* scratch, size_t scratch_bytes, int B,
                       int H, int W, int C, void* stream);

/* Backward of dfir_nonlocal: grad_x [B][H][W][64]; grad_w_tpg [24][64], grad_b_tpg [24], grad_w_out [64][8], grad_b_out
 * [64] are overwritten, or added to when accumulate != 0 (Q-SAN applies the one non-local block twice).  Max-pool
 * gradients go to the first maximum of each 2x2 window, as torch's max_pool2d backward does. */
size_t dfir_nonlocal_backward_scratch_bytes(int B, int H, int W);
int dfir_nonlocal_backward(const float* x, const float* grad_out, const float* w_tpg, const float* b_tpg,
                           const float* w_out, float* grad_x, float* grad_w_tpg, float* grad_b_tpg, float* grad_w_out,
                           float* grad_b_out, int accumulate, void* scratch, size_t scratch_bytes, int B, int H, int W,
                           int C, void* stream);

/* dfir_pack_conv3x3_f32 with the data-gradient option: transpose != 0 packs the 180-degree-rotated, channel-transposed
 * filter, so that dfir_conv3x3_f32(dy, packed, bias = NULL, ..., Cin = cout, Cout = cin) is the data gradient of the conv. */
int dfir_pack_conv3x3_f32_ex(const float* w_oihw, float* out, int cout, int cin, int transpose, void* stream);

/* Staged training step (Q-SAN / Q-HAN put their own layers between the stages of the Q-RCAN trunk; reference:
 * attention_manipulators/architectures.py:447-467, 514-540 under BaseModel.run_train, models/__init__.py:466-489).
 * One stage per call, feature maps between stages fp32 NHWC [B][H][W][C] owned by the caller, activations stashed in
 * the same workspace as dfir_qrcan_train_forward (dfir_qrcan_train_workspace_bytes):
 *   DFIR_TRAIN_HEAD   forward : x_nchw -> feat_out (also evaluates the meta-attention scales of every block);
 *                     backward: grad_feat_out -> head conv weight / bias gradients
 *   DFIR_TRAIN_GROUPS forward : feat_in -> groups [g_begin, g_end) -> feat_out (+ group_out[g - g_begin] after every
 *                     group, optional); backward: grad_feat_out -> grad_feat_in + the groups' conv gradients.  Networks
 *                     without group convs (Q-SAN) run one group per call
 *   DFIR_TRAIN_TAIL   forward : feat_in -> upsampler + tail conv -> out_nchw; backward: grad_out_nchw -> grad_feat_in
 *   DFIR_TRAIN_ATTN   backward only, after every GROUPS stage: channel- / meta-attention parameter gradients */
#define DFIR_TRAIN_HEAD 1
#define DFIR_TRAIN_GROUPS 2
#define DFIR_TRAIN_TAIL 8
#define DFIR_TRAIN_ATTN 16
int dfir_qrcan_train_stage_forward(const dfir_qrcan_net* net, int stage, int g_begin, int g_end, const float* x_nchw,
                                   const float* attributes, const float* feat_in, float* feat_out, float* group_out,
                                   float* out_nchw, int B, int H, int W, int precision, void* workspace,
                                   size_t workspace_bytes, void* stream);
int dfir_qrcan_train_stage_backward(const dfir_qrcan_net* net, const dfir_qrcan_params* grads, int stage, int g_begin,
                                    int g_end, const float* x_nchw, const float* attributes, const float* grad_out_nchw,
                                    const float* grad_feat_out, float* grad_feat_in, int B, int H, int W, int precision,
                                    void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DFIR_H_ */
