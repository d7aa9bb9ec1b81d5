// Internal declarations shared by the CUDA translation units (not part of the public C ABI — that is
// include/dfir.h).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/dfir.h"

namespace dfir {

// epilogues of the tensor-core conv
enum : int {
  EPI_BIAS = 0,       // out_bf16 = acc + b
  EPI_BIAS_RELU = 1,  // out_bf16 = relu(acc + b)                      (RCAB conv1, architectures.py:155-160)
  EPI_BIAS_POOL = 2,  // out_bf16 = acc + b ; per-row channel sums     (RCAB conv2 + CA avg-pool, :107)
  EPI_BIAS_SKIP = 3,  // out_f32 = acc + b + skip ; out_bf16 = bf16()  (group / trunk tail conv + `res += x`, :231-232)
  EPI_TAIL_NCHW = 4,  // out_f32 NCHW, cout<=16 real channels          (tail conv 64 -> 3, :303)
  EPI_RELU_STATS = 5, // RCAB conv1: out_bf16 = t = relu(acc + b) plus the sums of t that determine the channel
                      // attention of the block BEFORE its second conv runs (pool-by-linearity)
  EPI_SCALE_SKIP = 6, // RCAB conv2 / group conv: out_f32 = (acc + b) * s[b][c] + skip ; out_bf16 = bf16(out_f32)
                      // i.e. QCALayer/ParaCALayer `x * y` and `res += x` in the epilogue (:127,179; q_layer.py:43)
  EPI_RELU_MASK = 7,  // backward of conv-ReLU: out_bf16 = (acc + b) where the saved activation mask_bf16 > 0, else 0
  EPI_SCALE_SKIP_HL = 8,  // EPI_SCALE_SKIP of the inference chain on the bf16 hi + bf16 lo residual stream (x = hi + lo,
                          // 16 significant bits): out = (acc + b) * s + (skip_hi + skip_lo) -> out_hi = bf16(out),
                          // out_lo = bf16(out - out_hi).  10 instead of 12 bytes per element (the hi plane IS the next
                          // conv's operand); skip and result tiles travel by TMA, 16 pixels x 64 channels at a time.
  EPI_SCALE_SKIP_HL8 = 9, // the same with an 8-BIT lo plane: the stream value is a 24-bit float X (16 significant bits) whose
                          // bit pattern is (hi << 16) + (q << 8), hi = nearest bf16, q = int8: 8 bytes per element (t 2, hi
                          // in/out 4, lo in/out 2) and no more epilogue arithmetic than the bf16 lo plane needs
  EPI_RELU_STATS_W = 10,  // EPI_RELU_STATS with the image statistics in fixed point (istats) and warp-autonomous epilogue warps:
                          // private staging buffers and TMA stores of 32 pixels, per-thread 32-bit accumulators, no block barrier
};

enum : int { IN_TMA = 0, IN_FUSED = 1 };  // input modes of the tensor-core conv (see conv_tc.cu)

struct ConvTcArgs {  // kernel argument block
  int B, H, W, nseg, cin_off, cout;
  const void* wpacked;
  const float* bias;
  const float* skip_f32;
  float* out_f32;
  float* pool_rows;
  float* col_first;         // EPI_RELU_STATS: t at x = 0 / x = W-1 of every row, fp32 [B][H][64]
  float* col_last;
  const float* svec;        // EPI_SCALE_SKIP: per-image channel scale [B][64] (nullptr = 1)
  const __nv_bfloat16* mask_bf16;  // EPI_RELU_MASK: saved forward activation (dense NHWC)
  __nv_bfloat16* r_out;            // EPI_SCALE_SKIP (training forward): bf16 copy of acc + b, i.e. r before the scale
  float* ymean_out;                // EPI_SCALE_SKIP with epi_stats: pooled mean of r per image [B][64]
  int relu_out;                    // EPI_SCALE_SKIP: ReLU after the skip add (last K-chunk of a wide conv + ReLU)
  int tail_accumulate;             // EPI_TAIL_NCHW: add to what out_f32 already holds (K-chunks of a wide tail conv)
  __nv_bfloat16* out_bf16_direct;  // EPI_SCALE_SKIP writes its bf16 copy with plain coalesced stores
  // IN_FUSED: conv input = r * s[b] + xin (xout = fp32 copy of it for the rows the CTA owns), with
  // s[b] = CA_style(mean(r_b) from pool_rows, attributes[b]) * sq[b]   (style NONE: s = res_scale * sq)
  const __nv_bfloat16* r_bf16;
  const float* xin_f32;
  float* xout_f32;
  const float* ca_params;
  const float* attributes;
  const float* sq;
  float res_scale;
  int ca_style, ca_R, ca_M, ca_A;
  int epi_stats;            // EPI_SCALE_SKIP: evaluate the attention vector from the statistics of t in-kernel
  int debug_probe;
  // L2 eviction-priority policy words (ptx.cuh), all valid when use_hints != 0: TMA input rows, bf16 output,
  // fp32 skip rows, fp32 output
  int use_hints;
  unsigned long long pol_in, pol_out, pol_skip, pol_f32;
  const void* next_w;       // packed weights of the NEXT conv of the chain (or nullptr): pulled into L2 at kernel start
  int next_w_bytes;
  int hl_store_lo;          // EPI_SCALE_SKIP_HL: also write the lo plane (0 = the result only feeds a conv: hi suffices)
  // Image statistics of pool-by-linearity as 64-bit FIXED-POINT sums (2^-24 units) accumulated with atomics: integer
  // addition is associative, so the result does not depend on how rows are grouped into CTA bands — an image's statistics
  // (hence its output) stay bit-identical whatever else is in the batch — and conv2's prologue reads 9 x 64 numbers per
  // image instead of summing three per-row arrays.  Layout [B][9][64] long long: T, C0, CL, R0, RL, K00, K0W, KH0, KHW
  // (total, first / last column sums, first / last row sums, the four corner pixels).
  long long* istats;        // EPI_RELU_STATS: accumulate here instead of writing pool_rows / col_first / col_last;
                            // EPI_SCALE_SKIP_HL with epi_stats == 2: read the statistics from here
  long long* istats_clear;  // EPI_RELU_STATS: [B][9][64] buffer to zero for the NEXT block's conv1 (or nullptr)
  int flip;                 // EPI_SCALE_SKIP_HL: traverse images and rows in DESCENDING order (the rows the previous,
                            // ascending, kernel touched last are still in L2); requires W <= 128
};

// EPI_SCALE_SKIP_HL: tensor maps of the stream planes (box = 64 channels x 16 pixels, SWIZZLE_128B):
// 0 skip hi, 1 skip lo, 2 out hi, 3 out lo
struct ConvHlMaps {
  CUtensorMap m[4];
};

struct ConvTcDesc {  // host-side launch description
  int B, H, W;
  int cin_total, cin_off;  // channels of the input tensor / first channel of the 64-wide slice consumed
  int cout;                // real output channels (tail only)
  int epi;
  int in_mode;             // IN_TMA | IN_FUSED
  int num_sms;
  const void* in_bf16;
  long long in_pix_stride = 0, in_row_stride = 0, in_img_stride = 0;  // bytes; 0 = dense NHWC
  const void* mask_bf16 = nullptr;                                     // EPI_RELU_MASK
  void* r_out = nullptr;                                               // EPI_SCALE_SKIP: save r (training forward)
  float* ymean_out = nullptr;                                          // EPI_SCALE_SKIP + epi_stats: save mean(r)
  int relu_out = 0, tail_accumulate = 0;                               // see ConvTcArgs
  const void* wpacked;
  const float* bias;
  void* out_bf16;
  long long out_pix_stride, out_row_stride, out_img_stride;  // bytes (pixel-shuffle folds into these)
  const float* skip_f32;
  float* out_f32;
  const void* skip_hi = nullptr;  // EPI_SCALE_SKIP_HL: dense NHWC bf16 planes of the stream (out hi = out_bf16)
  const void* skip_lo = nullptr;
  void* out_lo = nullptr;         // nullptr: hi only
  int flip = 0;                   // EPI_SCALE_SKIP_HL: descending traversal (see ConvTcArgs::flip)
  long long* istats = nullptr;        // see ConvTcArgs
  long long* istats_clear = nullptr;
  float* pool_rows;
  float* col_first;
  float* col_last;
  const float* svec;
  const void* r_bf16;      // IN_FUSED
  const float* xin_f32;
  float* xout_f32;
  const float* ca_params;   // IN_FUSED: attention vector inputs (pool_rows above holds the pooled row sums)
  const float* attributes;
  const float* sq;
  float res_scale;
  int ca_style, ca_R, ca_M, ca_A;
  int epi_stats;
  const void* next_wpacked = nullptr;  // next conv's packed weights, prefetched into L2 (see ConvTcArgs::next_w)
};

int conv3x3_c64_tc(const ConvTcDesc& d, cudaStream_t stream);
int debug_watchdog(unsigned int* out8, int reset);
int debug_trace(unsigned long long* out1024);

// ---- SIMT kernels (simt.cu)
struct AttnParams {  // one RCAB's channel-attention parameters (fp32, device pointers into the packed blob)
  int style;         // DFIR_STYLE_*
  int C, R, M, A;    // channels, reduced channels, metadata size, attributes size (64 for modulate)
  const float* w[8]; // style-dependent list (see attn_vector() in simt.cu)
};

int pack_conv_weights_bf16(const float* w_oihw, void* out, int cout, int cin, int nt_rows, int co_begin,
                           int co_stride, cudaStream_t s);
int pack_conv_weights_f32(const float* w_oihw, float* out, int cout, int cin, cudaStream_t s);
int head_conv(const float* x_nchw, const float* w_packed, const float* bias, float* out_f32, __nv_bfloat16* out_bf16,
              int B, int Cin, int H, int W, int Cout, cudaStream_t s, __nv_bfloat16* out_lo = nullptr, int lo8 = 0);
int conv3x3_f32(const float* in, const float* w_packed, const float* bias, const float* skip, float* out, int B, int H,
                int W, int Cin, int Cout, int relu, int ps_r, int out_nchw, cudaStream_t s, const float* mask = nullptr);
// ---- training input pipeline (degrade.cu)
int batch_blur(const float* x, const float* kernels, int kernel_per_image, const float* noise, const float* sigma, float* out,
               int B, int C, int H, int W, int l, int clamp01, cudaStream_t s);
int pca_encode(const float* kernels, const float* pca, const float* sigma, float* code, int B, int n, int k, cudaStream_t s);
int postprocess_u8_blocks(int B, long long HW);  // partial sums per image the Y-PSNR pass needs (doubles)
int postprocess_u8(const float* x, const float* hr, unsigned char* rgb8, unsigned char* ycc8, float* y_psnr, double* scratch,
                   int B, long long HW, cudaStream_t s);
int pool_rows_f32(const float* in, float* pool_rows, int B, int H, int W, int C, cudaStream_t s);
int postprocess_rgb(const float* x, float* rgb, float* ycc, int B, long long HW, float lo, float hi, cudaStream_t s);
int meta_attention(const float* meta, const float* w1, const float* b1, const float* w2, const float* b2, float* out,
                   int nblk, int B, int M, int Hid, int C, int relu, const int* blk_enabled, float out_scale,
                   cudaStream_t s);
int scale_residual(const void* r, int r_is_bf16, const float* x_in, const float* pool_rows, int pool_nrows,
                   const AttnParams& ap, const float* attributes, const float* sq, float res_scale, float* x_out,
                   __nv_bfloat16* x_out_bf16, int B, int H, int W, int C, cudaStream_t s, float* y_out = nullptr,
                   const float* pa = nullptr);
int ca_from_stats(const float* pool_rows, const float* col_first, const float* col_last, const void* w2_packed,
                  const float* bias2, const AttnParams& ap, const float* attributes, const float* sq, float* svec, int B,
                  int H, int W, cudaStream_t s);
// ---- Q-HAN / Q-SAN layers (san_han.cu)
int stream_encode_hl8(const float* in, void* hi, void* lo8, long long n, cudaStream_t s);
int stream_decode_hl8(const void* hi, const void* lo8, float* out, long long n, cudaStream_t s);
int channel_scale(const float* x, const float* svec, const float* add, float alpha, float* out, int B, long long HW,
                  int C, cudaStream_t s);
int lam_forward(const float* stack, long long map_stride, float gamma, float* out, float* scratch, int N, int B, int HW,
                int C, cudaStream_t s);
size_t lam_scratch_floats(int B, int N);
int csam_forward(const float* x, const float* w27, float bias, float gamma, float* out, int B, int H, int W, int C,
                 cudaStream_t s);
int soca_forward(const float* x, const float* mlp, int R, float* svec, float* scratch, int B, int H, int W, int C,
                 cudaStream_t s);
size_t soca_scratch_floats(int B);
size_t covpool_scratch_floats(int B);
int covpool_forward(const float* x, float* cov, float* scratch, int B, int H, int W, int C, int crop1000, cudaStream_t s);
int covpool_backward(const float* x, const float* grad_cov, float* grad_x, float* scratch, int B, int H, int W, int C,
                     int crop1000, cudaStream_t s);
size_t sqrtm_scratch_floats(int B, int iters);
int sqrtm_forward(const float* cov, float* out, int B, int C, int iters, cudaStream_t s);
int sqrtm_backward(const float* cov, const float* grad_out, float* grad_in, float* scratch, int B, int C, int iters,
                   cudaStream_t s);
int nonlocal_forward(const float* x, const float* wq, const float* bq, const float* wW, const float* bW, float* out,
                     float* scratch, int B, int H, int W, int C, cudaStream_t s);
size_t nonlocal_scratch_floats(int B, int H, int W);
int nonlocal_project_pool(const float* x, const float* wq, const float* bq, float* proj, float* keys, int B, int H, int W,
                          cudaStream_t s);
// ---- backward of the Q-HAN / Q-SAN layers (san_han_bwd.cu)
size_t channel_dot_scratch_floats(int B, int C);
int channel_dot(const float* a, const float* b, float* out, float* total, int accumulate_total, float* scratch, int B,
                long long HW, int C, cudaStream_t s);
size_t outer_reduce_scratch_floats(int Ao, int Bi);
int outer_reduce(const float* A, int Ao, const float* Bm, int Bi, long long npix, float* out, float* colsum,
                 int accumulate, float* scratch, cudaStream_t s);
int soca_mlp_forward(const float* S, const float* mlp, int R, float* svec, int B, cudaStream_t s);
size_t soca_mlp_bwd_scratch_floats(int B, int R);
int soca_mlp_backward(const float* S, const float* dsvec, const float* mlp, int R, float* dS, float* dmlp, float* scratch,
                      int B, cudaStream_t s);
size_t lam_bwd_scratch_floats(int B, int N);
int lam_backward(const float* stack, long long map_stride, const float* att, float gamma, const float* dout, float* dstack,
                 long long dmap_stride, float* dgamma, float* scratch, int N, int B, int HW, int C, cudaStream_t s);
size_t csam_bwd_scratch_floats(int B, int H, int W, int C);
int csam_backward(const float* x, const float* dout, const float* w27, float bias, float gamma, float* dx, float* dw27,
                  float* dbias, float* dgamma, float* scratch, int B, int H, int W, int C, cudaStream_t s);
size_t nonlocal_bwd_scratch_floats(int B, int H, int W);
int nonlocal_backward(const float* x, const float* dz, const float* wq, const float* bq, const float* wW, float* dx,
                      float* dwq, float* dbq, float* dwW, float* dbW, int accumulate, float* scratch, int B, int H, int W,
                      int C, cudaStream_t s);
// ---- training step (train_kernels.cu, wgrad_tc.cu)
int pack_bf16_multi(const float* const* tbl, const float* direct, void* out, int n_tiles, int cout, int nt_rows,
                    int per_src, int transpose, cudaStream_t s, int j0 = 0);
int pack_f32_multi(const float* const* tbl, const float* direct, float* out, int n, int cout, int cin, int transpose,
                   cudaStream_t s);
int gather_strided(const float* const* tbl, const float* direct, int tbl_stride, int tbl_off, float* out, int n_rows,
                   int n, int per_src, int src_stride, long long out_stride, cudaStream_t s);
int bwd_reduce_chunks(int HW);
int bwd_reduce_ca(const float* g, const void* r, int r_is_bf16, float* part, unsigned int* tickets,
                  const float* pool_rows, int pool_nrows, const float* ymean, int HW, const AttnParams& ap,
                  const float* attributes, const float* sq, float out_scale, float* svec, float* dyv, float* sig,
                  int sig_stride, int B, cudaStream_t s);
int form_dr(const float* g, const float* svec, const float* dyv, void* dr, int dr_is_bf16, int B, int HW, int C,
            cudaStream_t s);
// pixel attention backward: du = gradient entering the channel-attention backward, part = per-CTA partial sums
int pa_backward_ctas(int B, int HW);
size_t pa_backward_part_floats(int B, int HW);
int pa_backward(const float* g, const void* r, int r_is_bf16, const float* ymean, const AttnParams& ap,
                const float* attributes, const float* sq, const float* pa, float* du, float* part, int B, int HW,
                cudaStream_t s);
int pa_finish(const float* part, const float* sq, float out_scale, float* sig, int sig_stride, int dzq_off,
              float* const* gpa, int B, int HW, cudaStream_t s);
int add_f32(const float* a, const float* b, float* out, void* out_bf16, long long n, cudaStream_t s);
int pixel_unshuffle_f32(const float* in, float* out, int B, int h, int w, int C, int r, cudaStream_t s);
int adam_flat(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, float wd,
              long long step, cudaStream_t s);
int wgrad_f32_chunks(int B, int H, int Cin, int Cout);
size_t wgrad_scratch_floats(int S, int Cin, int Cout);
int wgrad_f32(const float* dY, const float* X, float* scratch, int B, int H, int W, int Cin, int Cout, cudaStream_t s,
              int* S_out);
int wgrad_reduce(const float* part, const float* dbpart, int S, int Cin, int n_rows, float* const* w_tbl, int w_idx,
                 float* w_direct, float* const* b_tbl, int b_idx, float* b_direct, int co_begin, int co_stride,
                 cudaStream_t s, int co_major = 0);
int wgrad_c64_co_major();  // 1 when the selected 64-channel weight-gradient kernel writes [tap][co][ci] partials
size_t wgrad_small_scratch_floats(int B, int H, int C);
int wgrad_small(const float* I, const void* F, int f_is_bf16, float* scratch, int B, int H, int W, int C, int C3,
                int tail_mode, float* dw, float* db, cudaStream_t s);
int attn_param_grads(const float* sig, int sig_stride, const float* attributes, int A, const float* meta_w1,
                     const float* meta_b1, const float* meta_w2, const int* q_enabled, float* const* ca_g,
                     float* const* meta_g, int nblk, int B, int C, int R, int M, int Hid, int style, int meta_relu,
                     cudaStream_t s);
int wgrad_c64_grid(int B, int H, int W, int num_sms);
int wgrad_c64_tc(const void* dy, long long dy_pix, long long dy_row, long long dy_img, const void* x, float* scratch, int B,
                 int H, int W, int num_sms, cudaStream_t s, int* S_out);            // tcgen05 version (default)
int wgrad_tc_watchdog(unsigned int* out8, int reset);
int wgrad_c64(const void* dy, long long dy_pix, long long dy_row, long long dy_img, const void* x, float* scratch, int B,
              int H, int W, int num_sms, cudaStream_t s, int* S_out);
int nchw_to_nhwc_bf16(const float* in, __nv_bfloat16* out, int B, int C, int H, int W, cudaStream_t s);
int f32_to_bf16(const float* in, __nv_bfloat16* out, long long n, cudaStream_t s);

}  // namespace dfir
