// Training step of Q-RCAN / Q-EDSR on the B200 path: forward with an activation stash and the full backward,
// behind the C ABI of include/dfir.h ("training step" section).
//   reference: BaseModel.run_train / standard_update (/root/reference/Code/SISR/models/__init__.py:466-489) calling
//   autograd over QRCAN.forward (attention_manipulators/architectures.py:309-316).
//
// Backward of one RCAB  x' = conv2(relu(conv1(x))) * s + x,  s = CA(mean(r)) * meta_scale  (architectures.py:172-180):
//   ds      = sum_p g * r                       (bwd_reduce_gr, HBM-bound)
//   s, dyv  = attention-MLP backward            (ca_backward, one small CTA per image)
//   dr      = g * s + dyv                       (form_dr, HBM-bound; dyv = dL/d mean / HW is constant over pixels)
//   dW2,db2 = wgrad(dr, t);   dz = dgrad(dr, W2) masked by t > 0
//   dW1,db1 = wgrad(dz, x);   g  = dgrad(dz, W1) + g
// The data-gradient convs are the forward tensor-core kernel fed with transposed / rotated weights; the weight
// gradients are csrc/wgrad_tc.cu (bf16, tcgen05) or wgrad_f32 (parity mode).
#include "kernels.h"
#include "attn.cuh"

#include <algorithm>
#include <cstdlib>

using namespace dfir;

namespace {

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

struct Carver {
  uint8_t* base;
  size_t off;
  explicit Carver(void* p) : base(reinterpret_cast<uint8_t*>(p)), off(0) {}
  template <typename T>
  T* take(size_t bytes) {
    T* r = reinterpret_cast<T*>(base == nullptr ? nullptr : base + off);
    off += align256(bytes);
    return r;
  }
};

int up_stages(int scale, int* r) {
  if (scale == 3) { *r = 3; return 1; }
  if (scale == 2) { *r = 2; return 1; }
  if (scale == 4) { *r = 2; return 2; }
  if (scale == 8) { *r = 2; return 3; }
  *r = 0;
  return -1;
}

#define DFIR_TRY(expr)                \
  do {                                \
    int rc__ = (expr);                \
    if (rc__ != DFIR_OK) return rc__; \
  } while (0)

struct TrainWs {
  uint8_t* act; size_t slot_bytes;   // activation stash, operand format (bf16 on the tensor-core path, else fp32)
  uint8_t* U[3];                     // upsampler stage outputs, operand format
  float* pool; size_t pool_stride;   // [B*nseg*H*C] pooled row sums (scratch of the forward)
  float *colf, *coll;                // [B*H*C] first / last column of t (pool-by-linearity statistics)
  float* sq;                         // [nblk][B][C] meta-attention scales
  float *Hh, *XA, *XB;               // fp32 streams (tensor-core path only)
  float *G, *gs, *dF32;              // fp32 gradients of the group stream / block stream / trunk output
  uint8_t *DR, *DZ, *Gop, *DFop;     // operand-format gradients
  float* wpart; size_t wpart_floats; // weight-gradient partial sums
  float *red, *svec, *dyv, *sig; int sig_stride;
  float* ymean; unsigned int* tickets;   // [nblk][B][C] pooled means saved by the forward; per-image tickets
  float *PAdu, *papart;                  // pixel attention backward: gradient after the PALayer, per-CTA partial sums
  uint8_t* DUop[2]; float* DU32; float* DYF;
  float* small;
  size_t total;
};

// Stages of a training step.  dfir_qrcan_train_forward / _backward run all of them on the library's own buffers; the staged
// entry points (dfir_qrcan_train_stage_*) run ONE per call with the feature maps between stages owned by the caller
// (fp32 NHWC), so that Q-SAN / Q-HAN can put their own layers between the head, the residual groups and the tail.
enum : int { T_HEAD = 1, T_GROUPS = 2, T_TRUNK_CONV = 4, T_TAIL = 8, T_ATTN = 16, T_ALL = 31 };
struct TrainStage {
  int stages = T_ALL;
  int g0 = 0, g1 = 0;
  const float* feat_in = nullptr;    // forward: input feature map of the stage (GROUPS, TAIL)
  float* feat_out = nullptr;         // forward: output feature map of the stage (HEAD, GROUPS)
  float* group_out = nullptr;        // forward, GROUPS: [g1 - g0][B][H][W][C] stream after every group (optional)
  const float* gfeat_out = nullptr;  // backward: gradient of the stage's output feature map (HEAD, GROUPS)
  float* gfeat_in = nullptr;         // backward: gradient of the stage's input feature map (GROUPS, TAIL)
  bool all() const { return stages == T_ALL; }
};

bool train_supported(const dfir_qrcan_net* n, int precision, bool staged = false) {
  int r = 0;
  if (n == nullptr || up_stages(n->scale, &r) < 0) return false;
  if (n->n_groups < 1 || n->n_blocks < 1) return false;
  if (n->pa_blob != nullptr && (n->n_feats != 64 || n->style == DFIR_STYLE_NONE)) return false;
  if (n->no_group_conv && n->n_groups != 1 && !staged) return false;
  if (n->style < DFIR_STYLE_NONE || n->style > DFIR_STYLE_EXTENDED) return false;
  if (precision == DFIR_PREC_BF16_TC) return n->n_feats == 64;
  if (precision != DFIR_PREC_FP32_SIMT) return false;
  return n->n_feats % 64 == 0 && n->n_feats <= 256 && 256 % n->n_feats == 0;
}

TrainWs carve_train(const dfir_qrcan_net* n, int B, int H, int W, int precision, void* ws) {
  Carver c(ws);
  TrainWs w{};
  const bool tc = precision == DFIR_PREC_BF16_TC;
  const size_t elt = tc ? 2 : 4;
  const int C = n->n_feats;
  const size_t feat = static_cast<size_t>(B) * H * W * C;
  const int nblk = n->n_groups * n->n_blocks;
  const int nslots = 3 * nblk + n->n_groups + 2;
  int r = 0;
  const int nup = up_stages(n->scale, &r);
  const int nseg = (W + 127) / 128;
  w.slot_bytes = align256(feat * elt);
  w.act = c.take<uint8_t>(w.slot_bytes * nslots);
  size_t f = 1;
  for (int t = 0; t < nup; ++t) {
    f *= static_cast<size_t>(r) * r;
    w.U[t] = c.take<uint8_t>(feat * f * elt);
  }
  const size_t top = feat * f;  // elements of the largest upsampler output
  w.pool_stride = static_cast<size_t>(B) * nseg * H * C;
  w.pool = c.take<float>(w.pool_stride * 4);
  w.colf = c.take<float>(static_cast<size_t>(B) * H * C * 4);
  w.coll = c.take<float>(static_cast<size_t>(B) * H * C * 4);
  w.sq = c.take<float>(static_cast<size_t>(nblk) * B * C * 4);
  if (tc) {
    w.Hh = c.take<float>(feat * 4);
    w.XA = c.take<float>(feat * 4);
    w.XB = c.take<float>(feat * 4);
  }
  w.G = c.take<float>(feat * 4);
  w.gs = c.take<float>(feat * 4);
  w.dF32 = c.take<float>(feat * 4);
  w.DR = c.take<uint8_t>(feat * elt);
  w.DZ = c.take<uint8_t>(feat * elt);
  if (tc) {
    w.Gop = c.take<uint8_t>(feat * elt);
    w.DFop = c.take<uint8_t>(feat * elt);
  }
  if (tc) {
    w.wpart_floats = static_cast<size_t>(160) * (9 * 64 * 64 + 64);
  } else {
    const int s1 = wgrad_f32_chunks(B, H, C, C);
    size_t m = wgrad_scratch_floats(s1, C, C);
    int h = H;
    for (int t = 0; t < nup; ++t) {
      m = std::max(m, wgrad_scratch_floats(wgrad_f32_chunks(B, h, C, r * r * C), C, r * r * C));
      h *= r;
    }
    w.wpart_floats = m;
  }
  w.wpart = c.take<float>(w.wpart_floats * 4);
  w.red = c.take<float>(static_cast<size_t>(B) * 32 * C * 4);
  w.svec = c.take<float>(static_cast<size_t>(B) * C * 4);
  w.dyv = c.take<float>(static_cast<size_t>(B) * C * 4);
  w.sig_stride = make_attn_chain(n->style, C, std::max(1, n->reduced), n->num_metadata).sig_size;
  w.sig = c.take<float>(static_cast<size_t>(nblk) * B * w.sig_stride * 4);
  w.ymean = c.take<float>(static_cast<size_t>(nblk) * B * C * 4);
  w.tickets = c.take<unsigned int>(static_cast<size_t>(B) * 4);
  if (n->pa_blob != nullptr) {
    w.PAdu = c.take<float>(feat * 4);
    w.papart = c.take<float>(pa_backward_part_floats(B, H * W) * 4);
  }
  w.DUop[0] = c.take<uint8_t>(top * elt);
  w.DUop[1] = c.take<uint8_t>(top / (static_cast<size_t>(r) * r) * elt);
  if (tc) w.DU32 = c.take<float>(top / (static_cast<size_t>(r) * r) * 4);
  else w.DYF = c.take<float>(top * 4);
  w.small = c.take<float>(wgrad_small_scratch_floats(B, H * n->scale, C) * 4);
  w.total = c.off;
  return w;
}

AttnParams make_ap(const dfir_qrcan_net* n, int blk) {
  AttnParams ap{};
  ap.style = n->style;
  ap.C = n->n_feats;
  ap.R = std::max(1, n->reduced);
  ap.M = n->num_metadata;
  ap.A = n->attr_size;
  ap.w[0] = n->ca_blob + static_cast<size_t>(blk) * n->ca_stride;
  return ap;
}

struct Ctx {
  const dfir_qrcan_net* n;
  TrainWs w;
  int B, H, W, C, nb, ng, nblk, per_group, n_trunk, nup, r, nseg, sms, dzq_off;
  bool tc, has_ca;
  cudaStream_t st;
  uint8_t* slot(int i) const { return w.act + static_cast<size_t>(i) * w.slot_bytes; }
  uint8_t* XIN(int k) const { return slot(3 * k); }
  uint8_t* T(int k) const { return slot(3 * k + 1); }
  uint8_t* R(int k) const { return slot(3 * k + 2); }
  uint8_t* XLAST(int g) const { return slot(3 * nblk + g); }
  uint8_t* TRUNK_IN() const { return n->no_group_conv ? XLAST(0) : slot(3 * nblk + ng); }
  uint8_t* F() const { return slot(3 * nblk + ng + 1); }
  float* pool(int) const { return w.pool; }
  const float* sq(int k) const { return n->any_q ? w.sq + static_cast<size_t>(k) * B * C : nullptr; }
  float* sig(int k) const { return w.sig + static_cast<size_t>(k) * B * w.sig_stride; }
  float* ymean(int k) const { return w.ymean + static_cast<size_t>(k) * B * C; }
  const float* pa(int k) const { return n->pa_blob != nullptr ? n->pa_blob + static_cast<size_t>(k) * n->pa_stride : nullptr; }
};

int make_ctx(Ctx& c, const dfir_qrcan_net* n, int B, int H, int W, int precision, void* ws, size_t ws_bytes, void* stream,
             bool staged = false) {
  if (!train_supported(n, precision, staged) || B <= 0 || H <= 0 || W <= 0) return DFIR_ERR_ARG;
  if (dfir_check_device() != DFIR_OK) return DFIR_ERR_ARCH;
  c.n = n;
  c.w = carve_train(n, B, H, W, precision, ws);
  if (ws == nullptr || c.w.total > ws_bytes) return DFIR_ERR_WORKSPACE;
  c.B = B; c.H = H; c.W = W; c.C = n->n_feats; c.nb = n->n_blocks; c.ng = n->n_groups; c.nblk = c.nb * c.ng;
  c.per_group = 2 * c.nb + (n->no_group_conv ? 0 : 1);
  c.n_trunk = c.ng * c.per_group + 1;
  c.nup = up_stages(n->scale, &c.r);
  c.nseg = (W + 127) / 128;
  c.tc = precision == DFIR_PREC_BF16_TC;
  c.has_ca = n->style != DFIR_STYLE_NONE;
  c.dzq_off = make_attn_chain(n->style, n->n_feats, std::max(1, n->reduced), n->num_metadata).dzq_off;
  c.st = S(stream);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return DFIR_ERR_CUDA;
  c.sms = std::min(c.sms, 160);  // the weight-gradient scratch is sized for at most 160 CTAs
  if (c.tc && (n->conv_w_bf16 == nullptr || n->tail_w_bf16 == nullptr)) return DFIR_ERR_ARG;
  if (!c.tc && (n->conv_w_f32 == nullptr || n->up_w_f32 == nullptr || n->tail_w_f32 == nullptr)) return DFIR_ERR_ARG;
  return DFIR_OK;
}

// ---- tensor-core conv launch helper
ConvTcDesc tc_desc(const Ctx& c, const void* wpacked, const float* bias, int epi, int h, int w) {
  ConvTcDesc d{};
  d.B = c.B; d.H = h; d.W = w; d.cin_total = 64; d.cin_off = 0; d.cout = 64; d.epi = epi; d.in_mode = IN_TMA;
  d.num_sms = c.sms;
  d.wpacked = wpacked; d.bias = bias;
  d.out_pix_stride = 128; d.out_row_stride = static_cast<long long>(w) * 128;
  d.out_img_stride = static_cast<long long>(h) * w * 128;
  return d;
}
const uint8_t* tc_w(const Ctx& c, int widx) {
  return reinterpret_cast<const uint8_t*>(c.n->conv_w_bf16) + static_cast<size_t>(widx) * 9 * 64 * 128;
}
const uint8_t* tc_wT(const Ctx& c, int widx) {
  return reinterpret_cast<const uint8_t*>(c.n->conv_wT_bf16) + static_cast<size_t>(widx) * 9 * 64 * 128;
}
const float* tc_b(const Ctx& c, int widx) { return c.n->conv_b + static_cast<size_t>(widx) * 64; }

// block schedule of the training forward: pool-by-linearity (2 launches per RCAB) once a CTA's band is long enough to
// hide the attention prologue, else the streamer (3 launches)
bool train_linear_schedule(int B, int H, int W, int sms) {
  if (const char* e = getenv("DFIR_TRAIN_SCHEDULE")) {  // test hook: "linear" | "streamer"
    if (e[0] == 'l') return true;
    if (e[0] == 's') return false;
  }
  return static_cast<long long>(B) * ((W + 127) / 128) * H >= 16ll * std::max(1, sms);
}

// =============================================================================================== forward
int copy_f32(float* dst, const float* src, size_t n, cudaStream_t st) {
  return cudaMemcpyAsync(dst, src, n * 4, cudaMemcpyDeviceToDevice, st) == cudaSuccess ? DFIR_OK : DFIR_ERR_CUDA;
}

int train_forward_tc(const Ctx& c, const float* x, const float* attr, float* out, const TrainStage& ts) {
  const dfir_qrcan_net* n = c.n;
  const TrainWs& w = c.w;
  const int H = c.H, W = c.W, C = 64;
  const size_t feat = static_cast<size_t>(c.B) * H * W * C;
  if (ts.stages & T_HEAD) {
    if (ts.all())
      DFIR_TRY(head_conv(x, n->head_w_f32, n->head_b, w.Hh, reinterpret_cast<__nv_bfloat16*>(c.XIN(0)), c.B, n->in_feats,
                         H, W, C, c.st));
    else
      DFIR_TRY(head_conv(x, n->head_w_f32, n->head_b, ts.feat_out, nullptr, c.B, n->in_feats, H, W, C, c.st));
  }
  const bool linear = train_linear_schedule(c.B, H, W, c.sms) && n->pa_blob == nullptr;  // (PALayer lives in the streamer)
  const float* first_skip = w.Hh;
  const int g0 = (ts.stages & T_GROUPS) ? ts.g0 : 0, g1 = (ts.stages & T_GROUPS) ? ts.g1 : 0;
  if ((ts.stages & T_GROUPS) && !ts.all()) {
    DFIR_TRY(f32_to_bf16(ts.feat_in, reinterpret_cast<__nv_bfloat16*>(c.XIN(g0 * c.nb)), static_cast<long long>(feat), c.st));
    first_skip = ts.feat_in;
  }
  for (int g = g0; g < g1; ++g) {
    const float* skip32 = g == g0 ? first_skip : w.XA;
    for (int b = 0; b < c.nb; ++b) {
      const int k = g * c.nb + b;
      const int w1 = g * c.per_group + 2 * b, w2 = w1 + 1;
      uint8_t* next_bf = (b + 1 < c.nb) ? c.XIN(k + 1) : c.XLAST(g);
      if (!linear) {
        // streamer schedule: conv1, conv2 (+ pooled row sums), one bandwidth-shaped pass x' = r*s + x.  Measured at
        // 16 x 64x64 (7 rows per CTA): 12.8 ms per forward against 15.0 ms for the pool-by-linearity pair, whose
        // in-kernel attention prologue (~10 us) no longer hides behind the short pipeline.
        ConvTcDesc c1 = tc_desc(c, tc_w(c, w1), tc_b(c, w1), EPI_BIAS_RELU, H, W);
        c1.in_bf16 = c.XIN(k); c1.out_bf16 = c.T(k);
        DFIR_TRY(conv3x3_c64_tc(c1, c.st));
        ConvTcDesc c2 = tc_desc(c, tc_w(c, w2), tc_b(c, w2), c.has_ca ? EPI_BIAS_POOL : EPI_BIAS, H, W);
        c2.in_bf16 = c.T(k); c2.out_bf16 = c.R(k); c2.pool_rows = w.pool;
        DFIR_TRY(conv3x3_c64_tc(c2, c.st));
        DFIR_TRY(scale_residual(c.R(k), 1, b == 0 ? skip32 : w.XB, w.pool, c.nseg * H, make_ap(n, k), attr, c.sq(k), 1.f,
                                w.XB, reinterpret_cast<__nv_bfloat16*>(next_bf), c.B, H, W, C, c.st, c.ymean(k), c.pa(k)));
        continue;
      }
      // pool-by-linearity, as in inference: conv1 emits the statistics of t, conv2's prologue turns
      // them into the attention vector and its epilogue writes x' = r*s + x — plus, for the backward, r and mean(r)
      ConvTcDesc c1 = tc_desc(c, tc_w(c, w1), tc_b(c, w1), c.has_ca ? EPI_RELU_STATS : EPI_BIAS_RELU, H, W);
      c1.in_bf16 = c.XIN(k); c1.out_bf16 = c.T(k);
      c1.pool_rows = w.pool; c1.col_first = w.colf; c1.col_last = w.coll;
      DFIR_TRY(conv3x3_c64_tc(c1, c.st));
      ConvTcDesc c2 = tc_desc(c, tc_w(c, w2), tc_b(c, w2), EPI_SCALE_SKIP, H, W);
      c2.in_bf16 = c.T(k); c2.skip_f32 = b == 0 ? skip32 : w.XB; c2.out_f32 = w.XB; c2.out_bf16 = next_bf;
      c2.r_out = c.R(k);
      if (c.has_ca) {
        c2.pool_rows = w.pool; c2.col_first = w.colf; c2.col_last = w.coll; c2.epi_stats = 1;
        c2.ca_style = n->style; c2.ca_R = std::max(1, n->reduced); c2.ca_M = n->num_metadata; c2.ca_A = n->attr_size;
        c2.ca_params = n->ca_blob + static_cast<size_t>(k) * n->ca_stride; c2.attributes = attr; c2.sq = c.sq(k);
        c2.ymean_out = c.ymean(k);
      } else {
        c2.svec = c.sq(k);
      }
      DFIR_TRY(conv3x3_c64_tc(c2, c.st));
    }
    if (!n->no_group_conv) {
      const int wg = g * c.per_group + 2 * c.nb;
      ConvTcDesc ct = tc_desc(c, tc_w(c, wg), tc_b(c, wg), EPI_SCALE_SKIP, H, W);
      ct.in_bf16 = c.XLAST(g); ct.skip_f32 = skip32; ct.out_f32 = w.XA;
      ct.out_bf16 = (g + 1 < c.ng) ? c.XIN((g + 1) * c.nb) : c.TRUNK_IN();
      DFIR_TRY(conv3x3_c64_tc(ct, c.st));
    }
    if (ts.group_out != nullptr)
      DFIR_TRY(copy_f32(ts.group_out + static_cast<size_t>(g - g0) * feat, n->no_group_conv ? w.XB : w.XA, feat, c.st));
  }
  if ((ts.stages & T_GROUPS) && !ts.all() && ts.feat_out != nullptr && g1 > g0)
    DFIR_TRY(copy_f32(ts.feat_out, n->no_group_conv ? w.XB : w.XA, feat, c.st));
  if (ts.stages & T_TRUNK_CONV) {
    const int wf = c.ng * c.per_group;
    ConvTcDesc cf = tc_desc(c, tc_w(c, wf), tc_b(c, wf), EPI_SCALE_SKIP, H, W);
    cf.in_bf16 = c.TRUNK_IN(); cf.skip_f32 = w.Hh; cf.out_f32 = nullptr; cf.out_bf16 = c.F();
    DFIR_TRY(conv3x3_c64_tc(cf, c.st));
  }
  if (!(ts.stages & T_TAIL)) return DFIR_OK;
  if (!ts.all()) DFIR_TRY(f32_to_bf16(ts.feat_in, reinterpret_cast<__nv_bfloat16*>(c.F()), static_cast<long long>(feat), c.st));
  const void* cur = c.F();
  int h = H, wd = W;
  const int r = c.r;
  for (int t = 0; t < c.nup; ++t) {
    uint8_t* U = w.U[t];
    const long long oW = static_cast<long long>(wd) * r, oH = static_cast<long long>(h) * r;
    for (int s = 0; s < r * r; ++s) {
      const int i = s / r, j = s % r;
      const int widx = c.n_trunk + t * r * r + s;
      ConvTcDesc d = tc_desc(c, tc_w(c, widx), tc_b(c, widx), EPI_BIAS, h, wd);
      d.in_bf16 = cur;
      d.out_bf16 = U + (static_cast<long long>(i) * oW + j) * C * 2;
      d.out_pix_stride = static_cast<long long>(r) * C * 2;
      d.out_row_stride = static_cast<long long>(r) * oW * C * 2;
      d.out_img_stride = oH * oW * C * 2;
      DFIR_TRY(conv3x3_c64_tc(d, c.st));
    }
    cur = U;
    h *= r;
    wd *= r;
  }
  ConvTcDesc d = tc_desc(c, n->tail_w_bf16, n->tail_b, EPI_TAIL_NCHW, h, wd);
  d.cout = n->out_feats; d.in_bf16 = cur; d.out_f32 = out;
  return conv3x3_c64_tc(d, c.st);
}

int train_forward_f32(const Ctx& c, const float* x, const float* attr, float* out, const TrainStage& ts) {
  const dfir_qrcan_net* n = c.n;
  const TrainWs& w = c.w;
  const int H = c.H, W = c.W, C = c.C, B = c.B;
  const size_t wsz = static_cast<size_t>(9) * C * C;
  const size_t feat = static_cast<size_t>(B) * H * W * C;
  auto F32 = [](uint8_t* p) { return reinterpret_cast<float*>(p); };
  auto conv = [&](const float* in, int widx, int relu, const float* skip, float* o) {
    return conv3x3_f32(in, n->conv_w_f32 + widx * wsz, n->conv_b + static_cast<size_t>(widx) * C, skip, o, B, H, W, C, C,
                       relu, 1, 0, c.st);
  };
  if (ts.stages & T_HEAD)
    DFIR_TRY(head_conv(x, n->head_w_f32, n->head_b, ts.all() ? F32(c.XIN(0)) : ts.feat_out, nullptr, B, n->in_feats, H, W, C,
                       c.st));
  const int g0 = (ts.stages & T_GROUPS) ? ts.g0 : 0, g1 = (ts.stages & T_GROUPS) ? ts.g1 : 0;
  if ((ts.stages & T_GROUPS) && !ts.all()) DFIR_TRY(copy_f32(F32(c.XIN(g0 * c.nb)), ts.feat_in, feat, c.st));
  for (int g = g0; g < g1; ++g) {
    const float* gin = F32(c.XIN(g * c.nb));
    for (int b = 0; b < c.nb; ++b) {
      const int k = g * c.nb + b;
      const int w1 = g * c.per_group + 2 * b;
      DFIR_TRY(conv(F32(c.XIN(k)), w1, 1, nullptr, F32(c.T(k))));
      DFIR_TRY(conv(F32(c.T(k)), w1 + 1, 0, nullptr, F32(c.R(k))));
      if (c.has_ca) DFIR_TRY(pool_rows_f32(F32(c.R(k)), c.pool(k), B, H, W, C, c.st));
      float* next = (b + 1 < c.nb) ? F32(c.XIN(k + 1)) : F32(c.XLAST(g));
      DFIR_TRY(scale_residual(c.R(k), 0, F32(c.XIN(k)), c.pool(k), H, make_ap(n, k), attr, c.sq(k), 1.f, next, nullptr, B,
                              H, W, C, c.st, c.ymean(k), c.pa(k)));
    }
    const float* gres = F32(c.XLAST(g));
    if (!n->no_group_conv) {
      float* o = (g + 1 < c.ng) ? F32(c.XIN((g + 1) * c.nb)) : F32(c.TRUNK_IN());
      DFIR_TRY(conv(F32(c.XLAST(g)), g * c.per_group + 2 * c.nb, 0, gin, o));
      gres = o;
    }
    if (ts.group_out != nullptr) DFIR_TRY(copy_f32(ts.group_out + static_cast<size_t>(g - g0) * feat, gres, feat, c.st));
    if (!ts.all() && g + 1 == g1 && ts.feat_out != nullptr) DFIR_TRY(copy_f32(ts.feat_out, gres, feat, c.st));
  }
  if (ts.stages & T_TRUNK_CONV)
    DFIR_TRY(conv(F32(c.TRUNK_IN()), c.ng * c.per_group, 0, F32(c.XIN(0)), F32(c.F())));
  if (!(ts.stages & T_TAIL)) return DFIR_OK;
  if (!ts.all()) DFIR_TRY(copy_f32(F32(c.F()), ts.feat_in, feat, c.st));
  const float* cur = F32(c.F());
  int h = H, wd = W;
  const int r = c.r;
  size_t woff = 0, boff = 0;
  for (int t = 0; t < c.nup; ++t) {
    float* U = F32(w.U[t]);
    const int co = r * r * C;
    DFIR_TRY(conv3x3_f32(cur, n->up_w_f32 + woff, n->up_b + boff, nullptr, U, B, h, wd, C, co, 0, r, 0, c.st));
    woff += static_cast<size_t>(9) * C * co;
    boff += co;
    cur = U;
    h *= r;
    wd *= r;
  }
  return conv3x3_f32(cur, n->tail_w_f32, n->tail_b, nullptr, out, B, h, wd, C, n->out_feats, 0, 1, 1, c.st);
}

// =============================================================================================== backward
// weight + bias gradient of trunk conv `widx` (tensor-core path): dy, xin bf16 dense NHWC
int wgrad_tc(const Ctx& c, const dfir_qrcan_params* gr, const void* dy, const void* xin, int widx) {
  int S_ = 0;
  DFIR_TRY(wgrad_c64(dy, 0, 0, 0, xin, c.w.wpart, c.B, c.H, c.W, c.sms, c.st, &S_));
  return wgrad_reduce(c.w.wpart, c.w.wpart + static_cast<size_t>(S_) * 9 * 64 * 64, S_, 64, 64, gr->conv_w, widx, nullptr,
                      gr->conv_b, widx, nullptr, 0, 1, c.st, wgrad_c64_co_major());
}

int wgrad_f(const Ctx& c, const dfir_qrcan_params* gr, const float* dy, const float* xin, int widx) {
  int S_ = 0;
  const int C = c.C;
  DFIR_TRY(wgrad_f32(dy, xin, c.w.wpart, c.B, c.H, c.W, C, C, c.st, &S_));
  return wgrad_reduce(c.w.wpart, c.w.wpart + static_cast<size_t>(S_) * 9 * C * C, S_, C, C, gr->conv_w, widx, nullptr,
                      gr->conv_b, widx, nullptr, 0, 1, c.st);
}

int train_backward_tc(const Ctx& c, const dfir_qrcan_params* gr, const float* x, const float* attr, const float* gout,
                      const TrainStage& ts) {
  const dfir_qrcan_net* n = c.n;
  const TrainWs& w = c.w;
  const int H = c.H, W = c.W, C = 64, B = c.B, r = c.r;
  const int HW = H * W;
  const size_t feat = static_cast<size_t>(B) * HW * C;
  const float out_scale = n->style == DFIR_STYLE_NONE ? n->res_scale : 1.f;
  int h = H * n->scale, wd = W * n->scale;
  if (ts.stages & T_TAIL) {
  // ---- tail conv C -> out_feats
  DFIR_TRY(wgrad_small(gout, w.U[c.nup - 1], 1, w.small, B, h, wd, C, n->out_feats, 1, gr->tail_w, gr->tail_b, c.st));
  DFIR_TRY(head_conv(gout, n->tail_wT_f32, nullptr, nullptr, reinterpret_cast<__nv_bfloat16*>(w.DUop[0]), B, n->out_feats,
                     h, wd, C, c.st));
  // ---- upsampler stages, last to first
  uint8_t* cur = w.DUop[0];
  int pp = 0;
  for (int t = c.nup - 1; t >= 0; --t) {
    const int hh = h / r, ww = wd / r;
    const void* X = t == 0 ? c.F() : w.U[t - 1];
    float* dX32 = t == 0 ? w.dF32 : w.DU32;
    uint8_t* dXbf = t == 0 ? w.DFop : w.DUop[pp ^ 1];
    for (int s = 0; s < r * r; ++s) {
      const int i = s / r, j = s % r;
      const uint8_t* slice = cur + (static_cast<long long>(i) * wd + j) * C * 2;
      const long long ps = static_cast<long long>(r) * C * 2, rs = static_cast<long long>(r) * wd * C * 2;
      const long long is = static_cast<long long>(h) * wd * C * 2;
      int S_ = 0;
      DFIR_TRY(wgrad_c64(slice, ps, rs, is, X, w.wpart, B, hh, ww, c.sms, c.st, &S_));
      DFIR_TRY(wgrad_reduce(w.wpart, w.wpart + static_cast<size_t>(S_) * 9 * 64 * 64, S_, 64, 64, gr->up_w, t, nullptr,
                            gr->up_b, t, nullptr, s, r * r, c.st, wgrad_c64_co_major()));
      ConvTcDesc d = tc_desc(c, tc_wT(c, c.n_trunk + t * r * r + s), nullptr, EPI_SCALE_SKIP, hh, ww);
      d.in_bf16 = slice; d.in_pix_stride = ps; d.in_row_stride = rs; d.in_img_stride = is;
      d.skip_f32 = s == 0 ? nullptr : dX32; d.out_f32 = dX32; d.out_bf16 = dXbf;
      DFIR_TRY(conv3x3_c64_tc(d, c.st));
    }
    cur = dXbf;
    pp ^= 1;
    h = hh;
    wd = ww;
  }
  if (!ts.all()) DFIR_TRY(copy_f32(ts.gfeat_in, w.dF32, feat, c.st));
  }
  // ---- trunk tail conv: F = conv_f(trunk_in) + head
  if (ts.stages & T_TRUNK_CONV) {
    const int wf = c.ng * c.per_group;
    DFIR_TRY(wgrad_tc(c, gr, w.DFop, c.TRUNK_IN(), wf));
    ConvTcDesc d = tc_desc(c, tc_wT(c, wf), nullptr, EPI_SCALE_SKIP, H, W);
    d.in_bf16 = w.DFop; d.out_f32 = w.G; d.out_bf16 = w.Gop;
    DFIR_TRY(conv3x3_c64_tc(d, c.st));
  }
  const int g0 = (ts.stages & T_GROUPS) ? ts.g0 : 0, g1 = (ts.stages & T_GROUPS) ? ts.g1 : 0;
  if ((ts.stages & T_GROUPS) && !ts.all()) {
    DFIR_TRY(copy_f32(w.G, ts.gfeat_out, feat, c.st));
    DFIR_TRY(f32_to_bf16(w.G, reinterpret_cast<__nv_bfloat16*>(w.Gop), static_cast<long long>(feat), c.st));
  }
  for (int g = g1 - 1; g >= g0; --g) {
    float* gsp = n->no_group_conv ? w.G : w.gs;
    if (!n->no_group_conv) {
      const int wg = g * c.per_group + 2 * c.nb;
      DFIR_TRY(wgrad_tc(c, gr, w.Gop, c.XLAST(g), wg));
      ConvTcDesc d = tc_desc(c, tc_wT(c, wg), nullptr, EPI_SCALE_SKIP, H, W);
      d.in_bf16 = w.Gop; d.out_f32 = w.gs; d.out_bf16 = w.DR;
      DFIR_TRY(conv3x3_c64_tc(d, c.st));
    }
    for (int b = c.nb - 1; b >= 0; --b) {
      const int k = g * c.nb + b;
      const int w1 = g * c.per_group + 2 * b, w2 = w1 + 1;
      const float* gblk = gsp;  // gradient entering the block's scale (after the pixel attention backward, if any)
      if (c.pa(k) != nullptr) {
        DFIR_TRY(pa_backward(gsp, c.R(k), 1, c.ymean(k), make_ap(n, k), attr, c.sq(k), c.pa(k), w.PAdu, w.papart, B, HW, c.st));
        gblk = w.PAdu;
      }
      DFIR_TRY(bwd_reduce_ca(gblk, c.R(k), 1, w.red, w.tickets, c.pool(k), c.nseg * H, c.has_ca ? c.ymean(k) : nullptr, HW,
                             make_ap(n, k), attr, c.pa(k) != nullptr ? nullptr : c.sq(k), out_scale, w.svec, w.dyv, c.sig(k),
                             w.sig_stride, B, c.st));
      if (c.pa(k) != nullptr)
        DFIR_TRY(pa_finish(w.papart, c.sq(k), out_scale, c.sig(k), w.sig_stride, c.dzq_off, gr->pa + 4 * k, B, HW, c.st));
      DFIR_TRY(form_dr(gblk, w.svec, c.has_ca ? w.dyv : nullptr, w.DR, 1, B, HW, C, c.st));
      DFIR_TRY(wgrad_tc(c, gr, w.DR, c.T(k), w2));
      {
        ConvTcDesc d = tc_desc(c, tc_wT(c, w2), nullptr, EPI_RELU_MASK, H, W);
        d.in_bf16 = w.DR; d.mask_bf16 = c.T(k); d.out_bf16 = w.DZ;
        DFIR_TRY(conv3x3_c64_tc(d, c.st));
      }
      DFIR_TRY(wgrad_tc(c, gr, w.DZ, c.XIN(k), w1));
      {
        ConvTcDesc d = tc_desc(c, tc_wT(c, w1), nullptr, EPI_SCALE_SKIP, H, W);
        d.in_bf16 = w.DZ; d.skip_f32 = gsp; d.out_f32 = gsp; d.out_bf16 = w.DR;
        DFIR_TRY(conv3x3_c64_tc(d, c.st));
      }
    }
    if (!n->no_group_conv) DFIR_TRY(add_f32(w.G, w.gs, w.G, w.Gop, static_cast<long long>(B) * HW * C, c.st));
  }
  if ((ts.stages & T_GROUPS) && !ts.all()) DFIR_TRY(copy_f32(ts.gfeat_in, w.G, feat, c.st));
  // ---- head conv: dL/d head output = dL/dF (trunk skip) + dL/d (group 0 input)
  if (ts.stages & T_HEAD) {
    if (ts.all()) DFIR_TRY(add_f32(w.dF32, w.G, w.G, nullptr, static_cast<long long>(B) * HW * C, c.st));
    DFIR_TRY(wgrad_small(x, ts.all() ? w.G : ts.gfeat_out, 0, w.small, B, H, W, C, n->in_feats, 0, gr->head_w, gr->head_b,
                         c.st));
  }
  if (!(ts.stages & T_ATTN)) return DFIR_OK;
  return attn_param_grads(w.sig, w.sig_stride, attr, n->attr_size, n->meta_w1, n->meta_b1, n->meta_w2, n->q_enabled,
                          c.has_ca ? gr->ca : nullptr, n->any_q ? gr->meta : nullptr, c.nblk, B, C,
                          std::max(1, n->reduced), n->num_metadata, n->meta_hidden, n->style, n->meta_relu, c.st);
}

int train_backward_f32(const Ctx& c, const dfir_qrcan_params* gr, const float* x, const float* attr, const float* gout,
                       const TrainStage& ts) {
  const dfir_qrcan_net* n = c.n;
  const TrainWs& w = c.w;
  const int H = c.H, W = c.W, C = c.C, B = c.B, r = c.r;
  const int HW = H * W;
  const size_t feat = static_cast<size_t>(B) * HW * C;
  const size_t wsz = static_cast<size_t>(9) * C * C;
  const float out_scale = n->style == DFIR_STYLE_NONE ? n->res_scale : 1.f;
  auto F32 = [](uint8_t* p) { return reinterpret_cast<float*>(p); };
  auto dgrad = [&](const float* dy, int widx, const float* skip, const float* mask, float* o) {
    return conv3x3_f32(dy, n->conv_wT_f32 + widx * wsz, nullptr, skip, o, B, H, W, C, C, 0, 1, 0, c.st, mask);
  };
  int h = H * n->scale, wd = W * n->scale;
  if (ts.stages & T_TAIL) {
  DFIR_TRY(wgrad_small(gout, w.U[c.nup - 1], 0, w.small, B, h, wd, C, n->out_feats, 1, gr->tail_w, gr->tail_b, c.st));
  DFIR_TRY(head_conv(gout, n->tail_wT_f32, nullptr, F32(w.DUop[0]), nullptr, B, n->out_feats, h, wd, C, c.st));
  float* cur = F32(w.DUop[0]);
  int pp = 0;
  for (int t = c.nup - 1; t >= 0; --t) {
    const int hh = h / r, ww = wd / r, co = r * r * C;
    const float* X = t == 0 ? F32(c.F()) : F32(w.U[t - 1]);
    float* dX = t == 0 ? w.dF32 : F32(w.DUop[pp ^ 1]);
    DFIR_TRY(pixel_unshuffle_f32(cur, w.DYF, B, hh, ww, C, r, c.st));
    int S_ = 0;
    DFIR_TRY(wgrad_f32(w.DYF, X, w.wpart, B, hh, ww, C, co, c.st, &S_));
    DFIR_TRY(wgrad_reduce(w.wpart, w.wpart + static_cast<size_t>(S_) * 9 * C * co, S_, C, co, gr->up_w, t, nullptr,
                          gr->up_b, t, nullptr, 0, 1, c.st));
    DFIR_TRY(conv3x3_f32(w.DYF, n->up_wT_f32 + static_cast<size_t>(t) * 9 * co * C, nullptr, nullptr, dX, B, hh, ww, co, C,
                         0, 1, 0, c.st));
    cur = dX;
    pp ^= 1;
    h = hh;
    wd = ww;
  }
  if (!ts.all()) DFIR_TRY(copy_f32(ts.gfeat_in, w.dF32, feat, c.st));
  }
  if (ts.stages & T_TRUNK_CONV) {
    const int wf = c.ng * c.per_group;
    DFIR_TRY(wgrad_f(c, gr, w.dF32, F32(c.TRUNK_IN()), wf));
    DFIR_TRY(dgrad(w.dF32, wf, nullptr, nullptr, w.G));
  }
  const int g0 = (ts.stages & T_GROUPS) ? ts.g0 : 0, g1 = (ts.stages & T_GROUPS) ? ts.g1 : 0;
  if ((ts.stages & T_GROUPS) && !ts.all()) DFIR_TRY(copy_f32(w.G, ts.gfeat_out, feat, c.st));
  for (int g = g1 - 1; g >= g0; --g) {
    float* gsp = n->no_group_conv ? w.G : w.gs;
    if (!n->no_group_conv) {
      const int wg = g * c.per_group + 2 * c.nb;
      DFIR_TRY(wgrad_f(c, gr, w.G, F32(c.XLAST(g)), wg));
      DFIR_TRY(dgrad(w.G, wg, nullptr, nullptr, w.gs));
    }
    for (int b = c.nb - 1; b >= 0; --b) {
      const int k = g * c.nb + b;
      const int w1 = g * c.per_group + 2 * b, w2 = w1 + 1;
      const float* gblk = gsp;
      if (c.pa(k) != nullptr) {
        DFIR_TRY(pa_backward(gsp, c.R(k), 0, c.ymean(k), make_ap(n, k), attr, c.sq(k), c.pa(k), w.PAdu, w.papart, B, HW, c.st));
        gblk = w.PAdu;
      }
      DFIR_TRY(bwd_reduce_ca(gblk, c.R(k), 0, w.red, w.tickets, c.pool(k), H, c.has_ca ? c.ymean(k) : nullptr, HW,
                             make_ap(n, k), attr, c.pa(k) != nullptr ? nullptr : c.sq(k), out_scale, w.svec, w.dyv, c.sig(k),
                             w.sig_stride, B, c.st));
      if (c.pa(k) != nullptr)
        DFIR_TRY(pa_finish(w.papart, c.sq(k), out_scale, c.sig(k), w.sig_stride, c.dzq_off, gr->pa + 4 * k, B, HW, c.st));
      DFIR_TRY(form_dr(gblk, w.svec, c.has_ca ? w.dyv : nullptr, w.DR, 0, B, HW, C, c.st));
      DFIR_TRY(wgrad_f(c, gr, F32(w.DR), F32(c.T(k)), w2));
      DFIR_TRY(dgrad(F32(w.DR), w2, nullptr, F32(c.T(k)), F32(w.DZ)));
      DFIR_TRY(wgrad_f(c, gr, F32(w.DZ), F32(c.XIN(k)), w1));
      DFIR_TRY(dgrad(F32(w.DZ), w1, gsp, nullptr, gsp));
    }
    if (!n->no_group_conv) DFIR_TRY(add_f32(w.G, w.gs, w.G, nullptr, static_cast<long long>(B) * HW * C, c.st));
  }
  if ((ts.stages & T_GROUPS) && !ts.all()) DFIR_TRY(copy_f32(ts.gfeat_in, w.G, feat, c.st));
  if (ts.stages & T_HEAD) {
    if (ts.all()) DFIR_TRY(add_f32(w.dF32, w.G, w.G, nullptr, static_cast<long long>(B) * HW * C, c.st));
    DFIR_TRY(wgrad_small(x, ts.all() ? w.G : ts.gfeat_out, 0, w.small, B, H, W, C, n->in_feats, 0, gr->head_w, gr->head_b,
                         c.st));
  }
  if (!(ts.stages & T_ATTN)) return DFIR_OK;
  return attn_param_grads(w.sig, w.sig_stride, attr, n->attr_size, n->meta_w1, n->meta_b1, n->meta_w2, n->q_enabled,
                          c.has_ca ? gr->ca : nullptr, n->any_q ? gr->meta : nullptr, c.nblk, B, C,
                          std::max(1, n->reduced), n->num_metadata, n->meta_hidden, n->style, n->meta_relu, c.st);
}

int meta_scales(const Ctx& c, const float* attr) {
  const dfir_qrcan_net* n = c.n;
  if (!n->any_q) return DFIR_OK;
  return meta_attention(attr, n->meta_w1, n->meta_b1, n->meta_w2, n->meta_b2, c.w.sq, c.nblk, c.B, n->num_metadata,
                        n->meta_hidden, n->n_feats, n->meta_relu, n->q_enabled,
                        n->style == DFIR_STYLE_NONE ? n->res_scale : 1.f, c.st);
}

}  // namespace

extern "C" {

int dfir_qrcan_repack(const dfir_qrcan_net* n, const dfir_qrcan_params* p, int precision, int with_backward,
                      void* stream) {
  if (n == nullptr || p == nullptr) return DFIR_ERR_ARG;
  int r = 0;
  const int nup = up_stages(n->scale, &r);
  if (nup < 0) return DFIR_ERR_ARG;
  const bool tc = precision == DFIR_PREC_BF16_TC;
  if (tc && n->n_feats != 64) return DFIR_ERR_ARG;
  cudaStream_t st = S(stream);
  const int C = n->n_feats, nb = n->n_blocks, ng = n->n_groups, nblk = nb * ng;
  const int per_group = 2 * nb + (n->no_group_conv ? 0 : 1);
  const int n_trunk = ng * per_group + 1;
  const int rr = r * r;
  const float* const* cw = const_cast<const float* const*>(p->conv_w);
  const float* const* cb = const_cast<const float* const*>(p->conv_b);
  const float* const* uw = const_cast<const float* const*>(p->up_w);
  const float* const* ub = const_cast<const float* const*>(p->up_b);
  float* conv_b = const_cast<float*>(n->conv_b);
  // biases
  DFIR_TRY(gather_strided(cb, nullptr, 1, 0, conv_b, n_trunk, C, 1, 1, C, st));
  DFIR_TRY(gather_strided(nullptr, p->tail_b, 0, 0, const_cast<float*>(n->tail_b), 1, n->out_feats, 1, 1, 16, st));
  DFIR_TRY(gather_strided(nullptr, p->head_b, 0, 0, const_cast<float*>(n->head_b), 1, C, 1, 1, C, st));
  DFIR_TRY(pack_f32_multi(nullptr, p->head_w, const_cast<float*>(n->head_w_f32), 1, C, n->in_feats, 0, st));
  if (with_backward) {
    if (n->tail_wT_f32 == nullptr) return DFIR_ERR_ARG;
    DFIR_TRY(pack_f32_multi(nullptr, p->tail_w, const_cast<float*>(n->tail_wT_f32), 1, n->out_feats, C, 1, st));
  }
  if (tc) {
    const size_t wbytes = 9 * 64 * 128;
    uint8_t* wf = reinterpret_cast<uint8_t*>(const_cast<void*>(n->conv_w_bf16));
    DFIR_TRY(pack_bf16_multi(cw, nullptr, wf, n_trunk, 64, 64, 1, 0, st));
    for (int t = 0; t < nup; ++t) {
      DFIR_TRY(pack_bf16_multi(uw + t, nullptr, wf + (n_trunk + static_cast<size_t>(t) * rr) * wbytes, rr, rr * 64, 64, rr,
                               0, st));
      DFIR_TRY(gather_strided(ub + t, nullptr, 1, 0, conv_b + (n_trunk + static_cast<size_t>(t) * rr) * C, rr, C, rr, rr, C,
                              st));
    }
    DFIR_TRY(pack_bf16_multi(nullptr, p->tail_w, const_cast<void*>(n->tail_w_bf16), 1, n->out_feats, 16, 1, 0, st));
    if (with_backward) {
      if (n->conv_wT_bf16 == nullptr) return DFIR_ERR_ARG;
      uint8_t* wt = reinterpret_cast<uint8_t*>(const_cast<void*>(n->conv_wT_bf16));
      DFIR_TRY(pack_bf16_multi(cw, nullptr, wt, n_trunk, 64, 64, 1, 1, st));
      for (int t = 0; t < nup; ++t)
        DFIR_TRY(pack_bf16_multi(uw + t, nullptr, wt + (n_trunk + static_cast<size_t>(t) * rr) * wbytes, rr, rr * 64, 64, rr,
                                 1, st));
    }
  } else {
    const size_t usz = static_cast<size_t>(9) * C * rr * C;
    DFIR_TRY(pack_f32_multi(cw, nullptr, const_cast<float*>(n->conv_w_f32), n_trunk, C, C, 0, st));
    DFIR_TRY(pack_f32_multi(uw, nullptr, const_cast<float*>(n->up_w_f32), nup, rr * C, C, 0, st));
    DFIR_TRY(gather_strided(ub, nullptr, 1, 0, const_cast<float*>(n->up_b), nup, rr * C, 1, 1, rr * C, st));
    DFIR_TRY(pack_f32_multi(nullptr, p->tail_w, const_cast<float*>(n->tail_w_f32), 1, n->out_feats, C, 0, st));
    if (with_backward) {
      if (n->conv_wT_f32 == nullptr || n->up_wT_f32 == nullptr) return DFIR_ERR_ARG;
      DFIR_TRY(pack_f32_multi(cw, nullptr, const_cast<float*>(n->conv_wT_f32), n_trunk, C, C, 1, st));
      DFIR_TRY(pack_f32_multi(uw, nullptr, const_cast<float*>(n->up_wT_f32), nup, rr * C, C, 1, st));
    }
    (void)usz;
  }
  if (n->style != DFIR_STYLE_NONE) {
    if (p->ca == nullptr) return DFIR_ERR_ARG;
    const AttnChain ch = make_attn_chain(n->style, C, std::max(1, n->reduced), n->num_metadata);
    float* blob = const_cast<float*>(n->ca_blob);
    const float* const* ca = const_cast<const float* const*>(p->ca);
    for (int l = 0; l < ch.L; ++l) {  // table: 8 pointers per block = (W, b) of up to four chain layers
      const int kin = ch.nin[l] + (ch.cat[l] ? n->num_metadata : 0);
      DFIR_TRY(gather_strided(ca, nullptr, 8, 2 * l, blob + ch.woff[l], nblk, ch.nout[l] * kin, 1, 1, n->ca_stride, st));
      DFIR_TRY(gather_strided(ca, nullptr, 8, 2 * l + 1, blob + ch.boff[l], nblk, ch.nout[l], 1, 1, n->ca_stride, st));
    }
  }
  if (n->pa_blob != nullptr) {  // PALayer rows: W1[8][64] b1[8] W2[8] b2
    if (p->pa == nullptr || n->pa_stride < 529) return DFIR_ERR_ARG;
    const float* const* pt = const_cast<const float* const*>(p->pa);
    float* blob = const_cast<float*>(n->pa_blob);
    const int sizes[4] = {512, 8, 8, 1}, offs[4] = {0, 512, 520, 528};
    for (int i = 0; i < 4; ++i)
      DFIR_TRY(gather_strided(pt, nullptr, 4, i, blob + offs[i], nblk, sizes[i], 1, 1, n->pa_stride, st));
  }
  if (n->any_q && p->meta != nullptr) {  // (no table: every block is scaled by the constant out_scale, e.g. EDSR)
    const float* const* mt = const_cast<const float* const*>(p->meta);
    const int hid = n->meta_hidden, M = n->num_metadata;
    DFIR_TRY(gather_strided(mt, nullptr, 4, 0, const_cast<float*>(n->meta_w1), nblk, hid * M, 1, 1, static_cast<long long>(hid) * M, st));
    DFIR_TRY(gather_strided(mt, nullptr, 4, 1, const_cast<float*>(n->meta_b1), nblk, hid, 1, 1, hid, st));
    DFIR_TRY(gather_strided(mt, nullptr, 4, 2, const_cast<float*>(n->meta_w2), nblk, C * hid, 1, 1, static_cast<long long>(C) * hid, st));
    DFIR_TRY(gather_strided(mt, nullptr, 4, 3, const_cast<float*>(n->meta_b2), nblk, C, 1, 1, C, st));
  }
  return DFIR_OK;
}

size_t dfir_qrcan_train_workspace_bytes(const dfir_qrcan_net* net, int B, int H, int W, int precision) {
  if (!train_supported(net, precision, true) || B <= 0 || H <= 0 || W <= 0) return 0;
  return carve_train(net, B, H, W, precision, nullptr).total;
}

long long dfir_qrcan_train_launch_count(const dfir_qrcan_net* n, int B, int H, int W, int precision) {
  if (!train_supported(n, precision)) return 0;
  int r = 0;
  const long long nup = up_stages(n->scale, &r);
  const long long nblk = static_cast<long long>(n->n_groups) * n->n_blocks;
  const long long ngc = n->no_group_conv ? 0 : n->n_groups;
  const long long ca = n->style != DFIR_STYLE_NONE ? 1 : 0;
  if (precision == DFIR_PREC_BF16_TC) {
    const long long pa = n->pa_blob != nullptr ? 1 : 0;
    const long long fwd = 1 + (n->any_q ? 1 : 0) + nblk * ((train_linear_schedule(B, H, W, 148) && !pa) ? 2 : 3) + ngc + 1 + nup * r * r + 1;
    const long long bwd = 3 + nup * r * r * 3 + 3 + ngc * 4 + nblk * (8 + 2 * pa) + 1 + 2 + 1;
    return fwd + bwd;
  }
  const long long pa = n->pa_blob != nullptr ? 1 : 0;
  const long long fwd = 1 + (n->any_q ? 1 : 0) + nblk * (3 + ca) + ngc + 1 + nup + 1;
  const long long bwd = 3 + nup * 4 + 3 + ngc * 4 + nblk * (8 + 2 * pa) + 1 + 2 + 1;
  return fwd + bwd;
}

int dfir_qrcan_train_forward(const dfir_qrcan_net* net, const float* x, const float* attributes, float* out, int B, int H,
                             int W, int precision, void* workspace, size_t workspace_bytes, void* stream) {
  if (x == nullptr || attributes == nullptr || out == nullptr) return DFIR_ERR_ARG;
  Ctx c;
  DFIR_TRY(make_ctx(c, net, B, H, W, precision, workspace, workspace_bytes, stream));
  DFIR_TRY(meta_scales(c, attributes));
  TrainStage ts;
  ts.g1 = c.ng;
  return c.tc ? train_forward_tc(c, x, attributes, out, ts) : train_forward_f32(c, x, attributes, out, ts);
}

int dfir_qrcan_train_stage_forward(const dfir_qrcan_net* net, int stage, int g_begin, int g_end, const float* x,
                                   const float* attributes, const float* feat_in, float* feat_out, float* group_out,
                                   float* out, int B, int H, int W, int precision, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  if (stage != T_HEAD && stage != T_GROUPS && stage != T_TAIL) return DFIR_ERR_ARG;
  Ctx c;
  DFIR_TRY(make_ctx(c, net, B, H, W, precision, workspace, workspace_bytes, stream, true));
  TrainStage ts;
  ts.stages = stage; ts.g0 = g_begin; ts.g1 = g_end; ts.feat_in = feat_in; ts.feat_out = feat_out; ts.group_out = group_out;
  if (stage == T_HEAD) {
    if (x == nullptr || feat_out == nullptr || attributes == nullptr) return DFIR_ERR_ARG;
    DFIR_TRY(meta_scales(c, attributes));  // the scales of every block of the trunk, kept in the workspace
  } else if (stage == T_GROUPS) {
    if (feat_in == nullptr || attributes == nullptr || g_begin < 0 || g_end > c.ng || g_begin >= g_end) return DFIR_ERR_ARG;
    if (net->no_group_conv && g_end != g_begin + 1) return DFIR_ERR_ARG;  // such groups chain through the caller
  } else if (feat_in == nullptr || out == nullptr) {
    return DFIR_ERR_ARG;
  }
  return c.tc ? train_forward_tc(c, x, attributes, out, ts) : train_forward_f32(c, x, attributes, out, ts);
}

int dfir_qrcan_train_stage_backward(const dfir_qrcan_net* net, const dfir_qrcan_params* grads, int stage, int g_begin,
                                    int g_end, const float* x, const float* attributes, const float* grad_out,
                                    const float* grad_feat_out, float* grad_feat_in, int B, int H, int W, int precision,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  if (grads == nullptr || (stage != T_HEAD && stage != T_GROUPS && stage != T_TAIL && stage != T_ATTN)) return DFIR_ERR_ARG;
  Ctx c;
  DFIR_TRY(make_ctx(c, net, B, H, W, precision, workspace, workspace_bytes, stream, true));
  if (c.tc && (net->conv_wT_bf16 == nullptr || net->tail_wT_f32 == nullptr)) return DFIR_ERR_ARG;
  if (!c.tc && (net->conv_wT_f32 == nullptr || net->up_wT_f32 == nullptr || net->tail_wT_f32 == nullptr)) return DFIR_ERR_ARG;
  TrainStage ts;
  ts.stages = stage; ts.g0 = g_begin; ts.g1 = g_end; ts.gfeat_out = grad_feat_out; ts.gfeat_in = grad_feat_in;
  if (stage == T_HEAD) {
    if (x == nullptr || grad_feat_out == nullptr) return DFIR_ERR_ARG;
  } else if (stage == T_GROUPS) {
    if (grad_feat_out == nullptr || grad_feat_in == nullptr || attributes == nullptr || g_begin < 0 || g_end > c.ng ||
        g_begin >= g_end)
      return DFIR_ERR_ARG;
    if (net->no_group_conv && g_end != g_begin + 1) return DFIR_ERR_ARG;
    if (cudaMemsetAsync(c.w.tickets, 0, static_cast<size_t>(B) * 4, c.st) != cudaSuccess) return DFIR_ERR_CUDA;
  } else if (stage == T_TAIL) {
    if (grad_out == nullptr || grad_feat_in == nullptr) return DFIR_ERR_ARG;
  } else if (attributes == nullptr) {
    return DFIR_ERR_ARG;
  }
  return c.tc ? train_backward_tc(c, grads, x, attributes, grad_out, ts)
              : train_backward_f32(c, grads, x, attributes, grad_out, ts);
}

int dfir_qrcan_train_backward(const dfir_qrcan_net* net, const dfir_qrcan_params* grads, const float* x,
                              const float* attributes, const float* grad_out, int B, int H, int W, int precision,
                              void* workspace, size_t workspace_bytes, void* stream) {
  if (grads == nullptr || x == nullptr || attributes == nullptr || grad_out == nullptr) return DFIR_ERR_ARG;
  if (net != nullptr && net->pa_blob != nullptr && grads->pa == nullptr) return DFIR_ERR_ARG;
  Ctx c;
  DFIR_TRY(make_ctx(c, net, B, H, W, precision, workspace, workspace_bytes, stream));
  if (cudaMemsetAsync(c.w.tickets, 0, static_cast<size_t>(B) * 4, c.st) != cudaSuccess) return DFIR_ERR_CUDA;
  if (c.tc && (net->conv_wT_bf16 == nullptr || net->tail_wT_f32 == nullptr)) return DFIR_ERR_ARG;
  if (!c.tc && (net->conv_wT_f32 == nullptr || net->up_wT_f32 == nullptr || net->tail_wT_f32 == nullptr)) return DFIR_ERR_ARG;
  TrainStage ts;
  ts.g1 = c.ng;
  return c.tc ? train_backward_tc(c, grads, x, attributes, grad_out, ts)
              : train_backward_f32(c, grads, x, attributes, grad_out, ts);
}

size_t dfir_conv3x3_wgrad_scratch_bytes(int B, int H, int W, int Cin, int Cout, int precision) {
  if (precision == DFIR_PREC_BF16_TC) return static_cast<size_t>(160) * (9 * 64 * 64 + 64) * 4;
  return wgrad_scratch_floats(wgrad_f32_chunks(B, H, Cin, Cout), Cin, Cout) * 4;
}

int dfir_conv3x3_wgrad_c64(const void* dy, long long dps, long long drs, long long dis, const void* x, int B, int H, int W,
                           float* dw, float* db, int co_begin, int co_stride, void* scratch, size_t scratch_bytes,
                           void* stream) {
  if (dy == nullptr || x == nullptr || scratch == nullptr || co_stride < 1) return DFIR_ERR_ARG;
  if (scratch_bytes < dfir_conv3x3_wgrad_scratch_bytes(B, H, W, 64, 64, DFIR_PREC_BF16_TC)) return DFIR_ERR_WORKSPACE;
  DFIR_TRY(dfir_check_device());
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return DFIR_ERR_CUDA;
  int S_ = 0;
  float* sc = reinterpret_cast<float*>(scratch);
  DFIR_TRY(wgrad_c64(dy, dps, drs, dis, x, sc, B, H, W, std::min(sms, 160), S(stream), &S_));
#ifdef DFIR_PROBES
  if (getenv("DFIR_WGRAD_PROBE") != nullptr && (atoi(getenv("DFIR_WGRAD_PROBE")) & 8)) return DFIR_OK;  // timing only
#endif
  return wgrad_reduce(sc, sc + static_cast<size_t>(S_) * 9 * 64 * 64, S_, 64, 64, nullptr, 0, dw, nullptr, 0, db, co_begin,
                      co_stride, S(stream), wgrad_c64_co_major());
}

int dfir_conv3x3_wgrad_f32(const float* dy, const float* x, int B, int H, int W, int Cin, int Cout, float* dw, float* db,
                           void* scratch, size_t scratch_bytes, void* stream) {
  if (dy == nullptr || x == nullptr || scratch == nullptr || Cin % 4 != 0) return DFIR_ERR_ARG;
  if (scratch_bytes < dfir_conv3x3_wgrad_scratch_bytes(B, H, W, Cin, Cout, DFIR_PREC_FP32_SIMT)) return DFIR_ERR_WORKSPACE;
  int S_ = 0;
  float* sc = reinterpret_cast<float*>(scratch);
  DFIR_TRY(wgrad_f32(dy, x, sc, B, H, W, Cin, Cout, S(stream), &S_));
  return wgrad_reduce(sc, sc + static_cast<size_t>(S_) * 9 * Cin * Cout, S_, Cin, Cout, nullptr, 0, dw, nullptr, 0, db, 0, 1,
                      S(stream));
}

int dfir_conv3x3_c64_dgrad(const void* dy, long long dps, long long drs, long long dis, const void* wT, const void* mask,
                           const float* skip, float* out_f32, void* out_bf16, int B, int H, int W, void* stream) {
  if (dy == nullptr || wT == nullptr || out_bf16 == nullptr) return DFIR_ERR_ARG;
  DFIR_TRY(dfir_check_device());
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return DFIR_ERR_CUDA;
  ConvTcDesc d{};
  d.B = B; d.H = H; d.W = W; d.cin_total = 64; d.cin_off = 0; d.cout = 64; d.in_mode = IN_TMA; d.num_sms = sms;
  d.epi = mask != nullptr ? EPI_RELU_MASK : EPI_SCALE_SKIP;
  d.in_bf16 = dy; d.in_pix_stride = dps; d.in_row_stride = drs; d.in_img_stride = dis;
  d.wpacked = wT; d.bias = nullptr; d.mask_bf16 = mask; d.skip_f32 = skip; d.out_f32 = out_f32; d.out_bf16 = out_bf16;
  d.out_pix_stride = 128; d.out_row_stride = static_cast<long long>(W) * 128;
  d.out_img_stride = static_cast<long long>(H) * W * 128;
  return conv3x3_c64_tc(d, S(stream));
}

int dfir_pack_conv3x3_bf16_ex(const float* w, void* out, int cout, int nt_rows, int co_begin, int co_stride, int transpose,
                              void* stream) {
  if (w == nullptr || out == nullptr || (nt_rows != 64 && nt_rows != 16) || co_stride < 1 || co_begin < 0 ||
      co_begin >= co_stride)
    return DFIR_ERR_ARG;
  return pack_bf16_multi(nullptr, w, out, 1, cout, nt_rows, co_stride, transpose, S(stream), co_begin);
}

int dfir_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, long long step, void* stream) {
  if (params == nullptr || grads == nullptr || exp_avg == nullptr || exp_avg_sq == nullptr) return DFIR_ERR_ARG;
  return adam_flat(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, S(stream));
}

size_t dfir_conv3x3_wgrad_small_scratch_bytes(int B, int H, int C) { return wgrad_small_scratch_floats(B, H, C) * 4; }

int dfir_conv3x3_wgrad_small(const float* img, const void* feat, int feat_is_bf16, int B, int H, int W, int C, int C3,
                             int tail_mode, float* dw, float* db, void* scratch, size_t scratch_bytes, void* stream) {
  if (img == nullptr || feat == nullptr || dw == nullptr || db == nullptr || scratch == nullptr) return DFIR_ERR_ARG;
  if (scratch_bytes < dfir_conv3x3_wgrad_small_scratch_bytes(B, H, C)) return DFIR_ERR_WORKSPACE;
  return wgrad_small(img, feat, feat_is_bf16, reinterpret_cast<float*>(scratch), B, H, W, C, C3, tail_mode, dw, db,
                     S(stream));
}

}  // extern "C"
