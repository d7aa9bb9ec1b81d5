// C ABI of libdfir_b200.so (declared in include/dfir.h) and the whole-network forward schedules.
#include "kernels.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

using namespace dfir;

namespace {

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
constexpr int kDefaultSplit = 1;
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

// Images per pass.  Every trunk kernel carries ~12 us of fill/drain latency, so today fewer, larger passes
// beat L2-resident ones (profiles/r01_summary.md): a pass is only split to bound the workspace (~6 GiB).
int auto_chunk(int B, int H, int W, int precision) {
  const long long px = static_cast<long long>(H) * W;
  const long long ws_per_px = precision == DFIR_PREC_BF16_TC ? 4200 : 6500;  // incl. the upsampler buffers
  long long c = (6ll << 30) / std::max<long long>(1, px * ws_per_px);
  if (c < 1) c = 1;
  if (c > B) c = B;
  return static_cast<int>(c);
}

struct QrcanWs {
  float *Hh, *XA, *XB, *XB1, *pool, *sq, *colf, *coll, *svec;
  long long* istats[2];  // fixed-point image statistics of pool-by-linearity, double buffered by block parity
  __nv_bfloat16 *Hbf, *XAbf, *XBbf, *T, *R;
  float *T32, *R32;
  void* U[3];
  size_t total;
};

QrcanWs carve_qrcan(const dfir_qrcan_net* n, int B, int Bc, int H, int W, int precision, void* ws) {
  Carver c(ws);
  QrcanWs w{};
  const size_t px = static_cast<size_t>(Bc) * H * W;
  const int C = n->n_feats;
  w.Hh = c.take<float>(px * C * 4);
  w.XA = c.take<float>(px * C * 4);
  w.XB = c.take<float>(px * C * 4);
  const int nseg = (W + 127) / 128;
  w.pool = c.take<float>(static_cast<size_t>(Bc) * nseg * H * C * 4);
  w.sq = c.take<float>(static_cast<size_t>(n->n_groups) * n->n_blocks * B * C * 4);
  w.colf = c.take<float>(static_cast<size_t>(Bc) * H * C * 4);
  w.coll = c.take<float>(static_cast<size_t>(Bc) * H * C * 4);
  w.svec = c.take<float>(static_cast<size_t>(Bc) * C * 4);
  w.istats[0] = c.take<long long>(static_cast<size_t>(Bc) * 576 * 8);
  w.istats[1] = c.take<long long>(static_cast<size_t>(Bc) * 576 * 8);
  int r = 0;
  const int nup = up_stages(n->scale, &r);
  if (precision == DFIR_PREC_BF16_TC) {
    w.Hbf = c.take<__nv_bfloat16>(px * C * 2);
    w.XAbf = c.take<__nv_bfloat16>(px * C * 2);
    w.XBbf = c.take<__nv_bfloat16>(px * C * 2);
    w.XB1 = c.take<float>(px * C * 4);
    w.T = c.take<__nv_bfloat16>(px * C * 2);
    w.R = c.take<__nv_bfloat16>(px * C * 2);
    size_t f = 1;
    for (int t = 0; t < nup; ++t) {
      f *= static_cast<size_t>(r) * r;
      w.U[t] = c.take<uint8_t>(px * f * C * 2);
    }
  } else {
    w.T32 = c.take<float>(px * C * 4);
    w.R32 = c.take<float>(px * C * 4);
    size_t f = 1;
    for (int t = 0; t < nup; ++t) {
      f *= static_cast<size_t>(r) * r;
      w.U[t] = c.take<uint8_t>(px * f * C * 4);
    }
  }
  w.total = c.off;
  return w;
}

#define DFIR_TRY(expr)            \
  do {                            \
    int rc__ = (expr);            \
    if (rc__ != DFIR_OK) return rc__; \
  } while (0)

AttnParams make_ap(const dfir_qrcan_net* n, int blk) {
  AttnParams ap{};
  ap.style = n->style;
  ap.C = n->n_feats;
  ap.R = n->reduced;
  ap.M = n->num_metadata;
  ap.A = n->attr_size;
  ap.w[0] = n->ca_blob + static_cast<size_t>(blk) * n->ca_stride;
  return ap;
}

// stage mask of the forward schedules
enum : int { ST_HEAD = 1, ST_GROUPS = 2, ST_TRUNK_TAIL = 4, ST_UPSAMPLE = 8, ST_ALL = 15 };

struct StageArgs {
  int stages = ST_ALL;
  int g_begin = 0, g_end = 0;     // groups to run (ST_GROUPS)
  int from_xa = 0;                // the stream entering g_begin is in XA/XAbf (external features), not the head output
  float* group_out = nullptr;     // optional: fp32 NHWC copy of the stream after every executed group
  // concurrent sub-passes (see SplitCtx): the first trunk conv waits for `phase_wait`, `phase_record` is recorded right
  // after it, so that consecutive sub-passes start one kernel apart and stay in anti-phase
  cudaEvent_t phase_wait = nullptr, phase_record = nullptr;
};

// Concurrent sub-passes.  An RCAB alternates a tensor-pipe-bound kernel (conv1: 134 MB, 38.7 GFLOP per 32 images) with an
// HBM-bound one (conv2: 335 MB for the same FLOPs).  Run back to back on all SMs each leaves the other resource idle; run
// as `ways` independent sub-passes (disjoint images, private workspaces, grids of #SMs / ways CTAs, one stream each,
// staggered by one kernel) the conv1 of one sub-pass overlaps the conv2 of another.  The streams and events are created
// once per device and live for the life of the process; fork / join are event edges, so the pattern is graph-capturable.
struct SplitCtx {
  static constexpr int kMaxWays = 4;
  cudaStream_t side[kMaxWays - 1] = {};
  cudaEvent_t fork = nullptr, join[kMaxWays - 1] = {}, phase[kMaxWays - 1] = {};
  bool ok = false;
};

SplitCtx* split_ctx(int dev) {
  static std::mutex mu;
  static SplitCtx ctx[64];
  if (dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  SplitCtx& c = ctx[dev];
  if (!c.ok) {
    bool good = cudaEventCreateWithFlags(&c.fork, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < SplitCtx::kMaxWays - 1 && good; ++i)
      good = cudaStreamCreateWithFlags(&c.side[i], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&c.join[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&c.phase[i], cudaEventDisableTiming) == cudaSuccess;
    if (!good) return nullptr;
    c.ok = true;
  }
  return &c;
}

// Number of concurrent sub-passes of one pass of Bc images (DFIR_SPLIT overrides; 1 = off).
int split_ways(const dfir_qrcan_net* n, int Bc, int precision) {
  static const int env = getenv("DFIR_SPLIT") == nullptr ? 0 : atoi(getenv("DFIR_SPLIT"));
  int ways = env > 0 ? env : kDefaultSplit;
  if (precision != DFIR_PREC_BF16_TC || n->n_feats != 64) return 1;
  if (ways > SplitCtx::kMaxWays) ways = SplitCtx::kMaxWays;
  while (ways > 1 && Bc < 2 * ways) --ways;  // at least two images per sub-pass
  return ways < 1 ? 1 : ways;
}

int qrcan_forward_bf16(const dfir_qrcan_net* n, const float* x, const float* attr, float* out, int B, int Bc, int b0,
                       int H, int W, const QrcanWs& w, int num_sms, const StageArgs& sa, cudaStream_t st) {
  const int C = 64;
  const int nb = n->n_blocks, ng = n->n_groups;
  const int per_group = 2 * nb + (n->no_group_conv ? 0 : 1);
  const int n_trunk = ng * per_group + 1;
  const bool has_ca = n->style != DFIR_STYLE_NONE;
  const size_t wbytes = 9 * 64 * 128;
  const uint8_t* cw = reinterpret_cast<const uint8_t*>(n->conv_w_bf16);
  const long long pixB = C * 2, rowB = static_cast<long long>(W) * C * 2, imgB = rowB * H;

  // Residual-stream format of the default (pool-by-linearity) schedule when the whole network runs in one call: two bf16
  // planes, x = hi + lo (16 significant bits; the hi plane is the conv operand).  conv2 then moves 10 instead of 12 bytes
  // per element.  The lo planes live in the memory of the fp32 stream buffers.  Everything that exposes the fp32 stream
  // (staged execution, per-group copies, the other schedules) keeps the fp32 format; DFIR_STREAM=f32 forces it.
  static const bool hl_allowed = getenv("DFIR_STREAM") == nullptr || strcmp(getenv("DFIR_STREAM"), "f32") != 0;
  const int sched0 = n->pa_blob != nullptr ? 2 : n->schedule;
  const bool hl = hl_allowed && sched0 == 0 && sa.stages == ST_ALL && sa.group_out == nullptr && !sa.from_xa;
  // 8-bit lo plane (x = hi + q * 2^(e - 15)): conv2 moves 8 instead of 10 bytes per element; DFIR_LO8=0 keeps the bf16 lo plane
  static const int hl_lo8 = getenv("DFIR_LO8") == nullptr ? 1 : atoi(getenv("DFIR_LO8"));
  const int EPI_HL = hl_lo8 ? EPI_SCALE_SKIP_HL8 : EPI_SCALE_SKIP_HL;
  static const int hl_flip = getenv("DFIR_FLIP") == nullptr ? 1 : atoi(getenv("DFIR_FLIP"));
  static const int hl_fixed_stats = getenv("DFIR_FIXED_STATS") == nullptr ? 1 : atoi(getenv("DFIR_FIXED_STATS"));
  if (hl && hl_fixed_stats && (sa.stages & ST_GROUPS) &&
      cudaMemsetAsync(w.istats[0], 0, static_cast<size_t>(Bc) * 576 * 8 * 2, st) != cudaSuccess)
    return DFIR_ERR_CUDA;  // both buffers start at zero; afterwards every conv1 clears the buffer of the next block
  __nv_bfloat16* const Hlo = reinterpret_cast<__nv_bfloat16*>(w.Hh);
  __nv_bfloat16* const XAlo = reinterpret_cast<__nv_bfloat16*>(w.XA);
  __nv_bfloat16* const XBlo = reinterpret_cast<__nv_bfloat16*>(w.XB);
  if (sa.stages & ST_HEAD)
    DFIR_TRY(head_conv(x + static_cast<size_t>(b0) * n->in_feats * H * W, n->head_w_f32, n->head_b, hl ? nullptr : w.Hh, w.Hbf,
                       Bc, n->in_feats, H, W, C, st, hl ? Hlo : nullptr, hl_lo8));
  const size_t feat_bytes = static_cast<size_t>(Bc) * H * W * C * 4;

  // One launch description shared by all trunk convs; the lambdas below fill in what differs.
  auto base = [&](int widx, int epi) {
    ConvTcDesc d{};
    d.B = Bc; d.H = H; d.W = W; d.cin_total = 64; d.cin_off = 0; d.cout = 64; d.epi = epi; d.in_mode = IN_TMA;
    d.num_sms = num_sms;
    d.wpacked = cw + static_cast<size_t>(widx) * wbytes; d.bias = n->conv_b + static_cast<size_t>(widx) * 64;
    if (widx + 1 < n_trunk) d.next_wpacked = cw + static_cast<size_t>(widx + 1) * wbytes;  // trunk convs run in index order
    d.out_pix_stride = pixB; d.out_row_stride = rowB; d.out_img_stride = imgB;
    d.pool_rows = w.pool;
    return d;
  };
  const float* attr_c = attr + static_cast<size_t>(b0) * n->attr_size;
  const int nseg = (W + 127) / 128;
  // operand of a fused conv: x_{b} = r_{b-1} * s_{b-1} + x_{b-1}, with s_{b-1} evaluated in the kernel prologue
  auto fuse_in = [&](ConvTcDesc& d, int prev_blk, const float* xprev, float* xnew) {
    d.in_mode = IN_FUSED; d.r_bf16 = w.R; d.xin_f32 = xprev; d.xout_f32 = xnew;
    d.ca_style = n->style; d.ca_R = n->reduced; d.ca_M = n->num_metadata; d.ca_A = n->attr_size;
    d.ca_params = n->ca_blob + static_cast<size_t>(prev_blk) * n->ca_stride;
    d.attributes = attr_c; d.res_scale = 1.f;
    d.sq = n->any_q ? w.sq + (static_cast<size_t>(prev_blk) * B + b0) * C : nullptr;
  };
  // Three schedules for the block chain (net->schedule):
  //   0 (default) pool-by-linearity: conv1 also emits the sums of t that determine the block's channel
  //     attention, a tiny kernel turns them into s, and conv2's epilogue computes x <- (conv2(t)+b)*s + x
  //     directly: r never exists in memory and there is no elementwise pass.
  //   1 fused-in : `r*s + x` folded into the next conv's input path (dfir_conv3x3_c64_fused).
  //   2 streamer : separate bandwidth-shaped elementwise kernel (dfir_ca_scale_residual).
  const int sched = n->pa_blob != nullptr ? 2 : n->schedule;  // pixel attention lives in the streamer kernel
  auto pa_of = [&](int blk) { return n->pa_blob != nullptr ? n->pa_blob + static_cast<size_t>(blk) * n->pa_stride : nullptr; };

  const float* xcur = w.Hh;
  const int gb = (sa.stages & ST_GROUPS) ? sa.g_begin : 0, ge = (sa.stages & ST_GROUPS) ? sa.g_end : 0;
  bool first_trunk = true;
  auto trunk_conv = [&](const ConvTcDesc& d) -> int {
    if (first_trunk && sa.phase_wait != nullptr && cudaStreamWaitEvent(st, sa.phase_wait, 0) != cudaSuccess) return DFIR_ERR_CUDA;
    DFIR_TRY(conv3x3_c64_tc(d, st));
    if (first_trunk && sa.phase_record != nullptr && cudaEventRecord(sa.phase_record, st) != cudaSuccess) return DFIR_ERR_CUDA;
    first_trunk = false;
    return DFIR_OK;
  };
  for (int g = gb; g < ge; ++g) {
    const bool from_head = g == 0 && !sa.from_xa;
    const float* skip32 = from_head ? w.Hh : w.XA;         // group input (fp32 stream), kept for `res += x`
    const __nv_bfloat16* gin = from_head ? w.Hbf : w.XAbf; // its bf16 copy = operand of the first conv
    xcur = skip32;                                          // x_b: fp32 stream entering block b
    for (int b = 0; b < nb; ++b) {
      const int blk = g * nb + b;
      const float* sq = n->any_q ? w.sq + (static_cast<size_t>(blk) * B + b0) * C : nullptr;
      // conv1: t = relu(conv(x_b))
      ConvTcDesc c1 = base(g * per_group + 2 * b, ((sched == 0 || sched == 3) && has_ca) ? EPI_RELU_STATS : EPI_BIAS_RELU);
      c1.out_bf16 = w.T; c1.col_first = w.colf; c1.col_last = w.coll;
      const bool fx = hl && hl_fixed_stats && has_ca && sched == 0;  // statistics as fixed-point sums (see ConvTcArgs::istats)
      if (fx) {
        c1.istats = w.istats[blk & 1];
        c1.istats_clear = w.istats[(blk + 1) & 1];
        // DFIR_STATS_W=1: warp-autonomous epilogue (private staging buffers / TMA stores, per-thread accumulators) instead of
        // the group-synchronous epilogue of EPI_RELU_STATS (default until it is measured faster)
        static const int stats_w = getenv("DFIR_STATS_W") == nullptr ? 0 : atoi(getenv("DFIR_STATS_W"));
        const int lastpx = (W - 1) % 128;   // (see the host check in conv3x3_c64_tc: one column accumulator per thread)
        if (stats_w && !(lastpx < 32 && lastpx % 8 == 0)) c1.epi = EPI_RELU_STATS_W;
      }
      if (b == 0) {
        c1.in_bf16 = gin;
      } else if (sched == 1) {
        float* xnew = (b & 1) ? w.XB : w.XB1;
        fuse_in(c1, blk - 1, xcur, xnew);
        xcur = xnew;
      } else {
        c1.in_bf16 = w.XBbf;
      }
      DFIR_TRY(trunk_conv(c1));
      if (sched == 0 || sched == 3) {
        const int w2 = g * per_group + 2 * b + 1;
        // conv2 + attention scale + residual: x_{b+1} = (conv(t) + b) * s + x_b (fp32, in place after block 0);
        // s is evaluated from the statistics of t inside the kernel while its pipeline fills.
        ConvTcDesc c2 = base(w2, hl ? EPI_HL : EPI_SCALE_SKIP);
        c2.in_bf16 = w.T; c2.skip_f32 = b == 0 ? skip32 : w.XB; c2.out_f32 = w.XB; c2.out_bf16 = w.XBbf;
        if (hl) {
          c2.skip_hi = b == 0 ? gin : w.XBbf;
          c2.skip_lo = b == 0 ? (from_head ? Hlo : XAlo) : XBlo;
          c2.out_lo = XBlo;
          c2.flip = hl_flip;  // conv1 walked its band downwards: conv2 walks it upwards, through what is still in L2
        }
        if (has_ca && sched == 3) {
          // schedule 3: the attention vector comes from its own small kernel instead of conv2's prologue
          DFIR_TRY(ca_from_stats(w.pool, w.colf, w.coll, cw + static_cast<size_t>(w2) * wbytes, n->conv_b + static_cast<size_t>(w2) * 64,
                                 make_ap(n, blk), attr_c, sq, w.svec, Bc, H, W, st));
          c2.svec = w.svec;
        } else if (has_ca) {
          c2.col_first = w.colf; c2.col_last = w.coll; c2.epi_stats = fx ? 2 : 1;
          c2.istats = fx ? w.istats[blk & 1] : nullptr;
          c2.ca_style = n->style; c2.ca_R = n->reduced; c2.ca_M = n->num_metadata; c2.ca_A = n->attr_size;
          c2.ca_params = n->ca_blob + static_cast<size_t>(blk) * n->ca_stride; c2.attributes = attr_c; c2.sq = sq;
        } else {
          c2.svec = sq;  // ParamResBlock: s = res_scale * meta scale (meta_attention already folded res_scale in)
        }
        DFIR_TRY(conv3x3_c64_tc(c2, st));
        continue;
      }
      // conv2: r = conv(t) + per-row pooled sums (the avg-pool of the channel attention)
      ConvTcDesc c2 = base(g * per_group + 2 * b + 1, has_ca ? EPI_BIAS_POOL : EPI_BIAS);
      c2.in_bf16 = w.T; c2.out_bf16 = w.R;
      DFIR_TRY(conv3x3_c64_tc(c2, st));
      if (sched == 2) {
        // x_{b+1} = r * s + x_b  (fp32 stream, in place after the first block) + bf16 copy for the next conv
        DFIR_TRY(scale_residual(w.R, 1, b == 0 ? skip32 : w.XB, w.pool, nseg * H, make_ap(n, blk), attr_c, sq, 1.f,
                                w.XB, w.XBbf, Bc, H, W, C, st, nullptr, pa_of(blk)));
      }
    }
    if (n->no_group_conv) {  // Q-EDSR / Q-SAN groups: the chain of blocks has no tail conv of its own
      if (sa.group_out != nullptr && sched != 1 && nb > 0 &&
          cudaMemcpyAsync(sa.group_out + static_cast<size_t>(g - gb) * Bc * H * W * C, w.XB, feat_bytes,
                          cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        return DFIR_ERR_CUDA;
      continue;
    }
    // group tail conv + `res += x` (group input)
    ConvTcDesc ct = base(g * per_group + 2 * nb, hl ? EPI_HL : EPI_SCALE_SKIP);
    ct.out_bf16 = w.XAbf; ct.skip_f32 = skip32; ct.out_f32 = w.XA; ct.svec = nullptr;
    if (hl) {
      ct.skip_hi = gin;
      ct.skip_lo = from_head ? Hlo : XAlo;
      ct.out_lo = XAlo;
    }
    if (nb == 0) {
      ct.in_bf16 = gin;
    } else if (sched == 1) {
      fuse_in(ct, g * nb + nb - 1, xcur, nullptr);
    } else {
      ct.in_bf16 = w.XBbf;
    }
    DFIR_TRY(conv3x3_c64_tc(ct, st));
    if (sa.group_out != nullptr &&
        cudaMemcpyAsync(sa.group_out + static_cast<size_t>(g - gb) * Bc * H * W * C, w.XA, feat_bytes,
                        cudaMemcpyDeviceToDevice, st) != cudaSuccess)
      return DFIR_ERR_CUDA;
  }
  __nv_bfloat16* trunk_out = w.XBbf;
  if (sa.stages & ST_TRUNK_TAIL) {
    ConvTcDesc cf = base(ng * per_group, hl ? EPI_HL : EPI_SCALE_SKIP);
    if (hl) {
      cf.skip_hi = w.Hbf;
      cf.skip_lo = Hlo;
      cf.out_lo = nullptr;  // the trunk output only feeds the upsampler convs
    }
    cf.in_bf16 = (ng == 0 || (n->no_group_conv && nb == 0)) ? w.Hbf : (n->no_group_conv ? w.XBbf : w.XAbf);
    if (sched == 1 && n->no_group_conv && ng > 0 && nb > 0) fuse_in(cf, ng * nb - 1, xcur, nullptr);  // last x never materialised
    // never in place: a band's first output row is another band's halo row
    trunk_out = cf.in_bf16 == w.XBbf ? w.XAbf : w.XBbf;
    cf.out_bf16 = trunk_out; cf.skip_f32 = w.Hh; cf.out_f32 = nullptr;
    DFIR_TRY(conv3x3_c64_tc(cf, st));
  }
  if (!(sa.stages & ST_UPSAMPLE)) return DFIR_OK;
  // upsampler: conv C -> r*r*C with PixelShuffle(r) folded into the TMA store strides
  int r = 0;
  const int nup = up_stages(n->scale, &r);
  const void* cur = trunk_out;
  int h = H, wd = W;
  for (int t = 0; t < nup; ++t) {
    uint8_t* U = reinterpret_cast<uint8_t*>(w.U[t]);
    const long long oW = static_cast<long long>(wd) * r, oH = static_cast<long long>(h) * r;
    for (int s = 0; s < r * r; ++s) {
      const int i = s / r, j = s % r;
      ConvTcDesc d{};
      const int widx = n_trunk + t * r * r + s;
      d.B = Bc; d.H = h; d.W = wd; d.cin_total = 64; d.cin_off = 0; d.cout = 64; d.epi = EPI_BIAS; d.num_sms = num_sms;
      d.in_mode = IN_TMA;
      d.in_bf16 = cur; d.wpacked = cw + static_cast<size_t>(widx) * wbytes; d.bias = n->conv_b + static_cast<size_t>(widx) * 64;
      d.out_bf16 = U + (static_cast<long long>(i) * oW + j) * C * 2;
      d.out_pix_stride = static_cast<long long>(r) * C * 2;
      d.out_row_stride = static_cast<long long>(r) * oW * C * 2;
      d.out_img_stride = oH * oW * C * 2;
      DFIR_TRY(conv3x3_c64_tc(d, st));
    }
    cur = U;
    h *= r;
    wd *= r;
  }
  {
    ConvTcDesc d{};
    d.B = Bc; d.H = h; d.W = wd; d.cin_total = 64; d.cin_off = 0; d.cout = n->out_feats; d.epi = EPI_TAIL_NCHW;
    d.num_sms = num_sms;
    d.in_bf16 = cur; d.wpacked = n->tail_w_bf16; d.bias = n->tail_b;
    d.out_f32 = out + static_cast<size_t>(b0) * n->out_feats * h * wd;
    DFIR_TRY(conv3x3_c64_tc(d, st));
  }
  return DFIR_OK;
}

int qrcan_forward_f32(const dfir_qrcan_net* n, const float* x, const float* attr, float* out, int B, int Bc, int b0,
                      int H, int W, const QrcanWs& w, const StageArgs& sa, cudaStream_t st) {
  const int C = n->n_feats;
  const int nb = n->n_blocks, ng = n->n_groups;
  const int per_group = 2 * nb + (n->no_group_conv ? 0 : 1);
  const size_t wsz = static_cast<size_t>(9) * C * C;
  if (sa.stages & ST_HEAD)
    DFIR_TRY(head_conv(x + static_cast<size_t>(b0) * n->in_feats * H * W, n->head_w_f32, n->head_b, w.Hh, nullptr, Bc,
                       n->in_feats, H, W, C, st));
  const size_t feat_bytes = static_cast<size_t>(Bc) * H * W * C * 4;
  const int gb = (sa.stages & ST_GROUPS) ? sa.g_begin : 0, ge = (sa.stages & ST_GROUPS) ? sa.g_end : 0;
  auto conv = [&](const float* in, int widx, int relu, const float* skip, float* o) {
    return conv3x3_f32(in, n->conv_w_f32 + widx * wsz, n->conv_b + static_cast<size_t>(widx) * C, skip, o, Bc, H, W, C,
                       C, relu, 1, 0, st);
  };
  for (int g = gb; g < ge; ++g) {
    const float* skip32 = (g == 0 && !sa.from_xa) ? w.Hh : w.XA;
    for (int b = 0; b < nb; ++b) {
      const int blk = g * nb + b;
      const float* cin = b == 0 ? skip32 : w.XB;
      DFIR_TRY(conv(cin, g * per_group + 2 * b, 1, nullptr, w.T32));
      DFIR_TRY(conv(w.T32, g * per_group + 2 * b + 1, 0, nullptr, w.R32));
      if (n->style != DFIR_STYLE_NONE) DFIR_TRY(pool_rows_f32(w.R32, w.pool, Bc, H, W, C, st));
      const float* sq = n->any_q ? w.sq + (static_cast<size_t>(blk) * B + b0) * C : nullptr;
      DFIR_TRY(scale_residual(w.R32, 0, cin, w.pool, H, make_ap(n, blk), attr + static_cast<size_t>(b0) * n->attr_size,
                              sq, 1.f, w.XB, nullptr, Bc, H, W, C, st, nullptr,
                              n->pa_blob != nullptr ? n->pa_blob + static_cast<size_t>(blk) * n->pa_stride : nullptr));
    }
    if (n->no_group_conv) {
      if (sa.group_out != nullptr && nb > 0 &&
          cudaMemcpyAsync(sa.group_out + static_cast<size_t>(g - gb) * Bc * H * W * C, w.XB, feat_bytes,
                          cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        return DFIR_ERR_CUDA;
      continue;
    }
    const float* cin = nb == 0 ? skip32 : w.XB;
    // out = conv(cin) + skip32 ; written to T32 first because XA may be the skip being read
    DFIR_TRY(conv(cin, g * per_group + 2 * nb, 0, skip32, w.T32));
    DFIR_TRY(cudaMemcpyAsync(w.XA, w.T32, static_cast<size_t>(Bc) * H * W * C * 4, cudaMemcpyDeviceToDevice, st) ==
                     cudaSuccess
                 ? DFIR_OK
                 : DFIR_ERR_CUDA);
    if (sa.group_out != nullptr &&
        cudaMemcpyAsync(sa.group_out + static_cast<size_t>(g - gb) * Bc * H * W * C, w.XA, feat_bytes,
                        cudaMemcpyDeviceToDevice, st) != cudaSuccess)
      return DFIR_ERR_CUDA;
  }
  if (sa.stages & ST_TRUNK_TAIL) {
    const float* fin = (ng == 0 || (n->no_group_conv && nb == 0)) ? w.Hh : (n->no_group_conv ? w.XB : w.XA);
    // XB may be the operand: write the trunk tail output to R32, which the upsampler then reads
    DFIR_TRY(conv(fin, ng * per_group, 0, w.Hh, w.R32));
  }
  if (!(sa.stages & ST_UPSAMPLE)) return DFIR_OK;
  int r = 0;
  const int nup = up_stages(n->scale, &r);
  const float* cur = w.R32;
  int h = H, wd = W;
  size_t woff = 0, boff = 0;
  for (int t = 0; t < nup; ++t) {
    float* U = reinterpret_cast<float*>(w.U[t]);
    const int co = r * r * C;
    DFIR_TRY(conv3x3_f32(cur, n->up_w_f32 + woff, n->up_b + boff, nullptr, U, Bc, h, wd, C, co, 0, r, 0, st));
    woff += static_cast<size_t>(9) * C * co;
    boff += co;
    cur = U;
    h *= r;
    wd *= r;
  }
  DFIR_TRY(conv3x3_f32(cur, n->tail_w_f32, n->tail_b, nullptr, out + static_cast<size_t>(b0) * n->out_feats * h * wd,
                       Bc, h, wd, C, n->out_feats, 0, 1, 1, st));
  return DFIR_OK;
}

}  // namespace

extern "C" {

const char* dfir_version(void) { return "dfir-b200 0.1 (abi 1, sm_100a)"; }

const char* dfir_error_string(int code) {
  switch (code) {
    case DFIR_OK: return "ok";
    case DFIR_ERR_ARG: return "invalid argument or unsupported configuration";
    case DFIR_ERR_CUDA: return "CUDA runtime error";
    case DFIR_ERR_DRIVER: return "cuTensorMapEncodeTiled driver entry point unavailable";
    case DFIR_ERR_TMAP: return "tensor map encoding failed";
    case DFIR_ERR_WORKSPACE: return "workspace too small";
    case DFIR_ERR_ARCH: return "device is not compute capability 10.x (sm_100a required)";
    default: return "unknown error";
  }
}

int dfir_debug_watchdog(unsigned int* out8_host, int reset) { return debug_watchdog(out8_host, reset); }
int dfir_debug_trace(unsigned long long* out1024_host) { return debug_trace(out1024_host); }

int dfir_check_device(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return DFIR_ERR_CUDA;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return DFIR_ERR_CUDA;
  return major == 10 ? DFIR_OK : DFIR_ERR_ARCH;
}

int dfir_pack_conv3x3_bf16(const float* w_oihw, void* out, int cout, int cin, int nt_rows, int co_begin,
                           int co_stride, void* stream) {
  return pack_conv_weights_bf16(w_oihw, out, cout, cin, nt_rows, co_begin, co_stride, S(stream));
}

int dfir_pack_conv3x3_f32(const float* w_oihw, float* out, int cout, int cin, void* stream) {
  return pack_conv_weights_f32(w_oihw, out, cout, cin, S(stream));
}

int dfir_conv3x3_c64(const void* in_bf16, int cin_total, int cin_off, const void* wpacked, const float* bias, int B,
                     int H, int W, int epi, int cout, void* out_bf16, long long out_pix_stride,
                     long long out_row_stride, long long out_img_stride, const float* skip_f32, float* out_f32,
                     float* pool_rows, int desc_mode, void* stream) {
  ConvTcDesc d{};
  d.B = B; d.H = H; d.W = W; d.cin_total = cin_total; d.cin_off = cin_off; d.cout = cout; d.epi = epi;
  d.in_mode = IN_TMA; d.num_sms = 0;
  d.in_bf16 = in_bf16; d.wpacked = wpacked; d.bias = bias; d.out_bf16 = out_bf16;
  d.out_pix_stride = out_pix_stride; d.out_row_stride = out_row_stride; d.out_img_stride = out_img_stride;
  d.skip_f32 = skip_f32; d.out_f32 = out_f32; d.pool_rows = pool_rows;
  if (desc_mode != 0) return DFIR_ERR_ARG;  // reserved (hardware bring-up variants were removed)
  if (epi == EPI_BIAS_POOL && pool_rows == nullptr) return DFIR_ERR_ARG;
  if (epi == EPI_BIAS_SKIP && skip_f32 == nullptr) return DFIR_ERR_ARG;
  if (epi == EPI_TAIL_NCHW && (out_f32 == nullptr || cout > 16)) return DFIR_ERR_ARG;
  if (epi != EPI_TAIL_NCHW && out_bf16 == nullptr) return DFIR_ERR_ARG;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return DFIR_ERR_CUDA;
  d.num_sms = sms;
  return conv3x3_c64_tc(d, S(stream));
}

int dfir_conv3x3_c64_fused(const void* r_bf16, const float* x_in, float* x_out, const float* pool_rows, int style,
                           const float* ca_params, int R, int M, int A, const float* attributes, const float* sq,
                           float res_scale, const void* wpacked, const float* bias, int B, int H, int W, int epi,
                           void* out_bf16, const float* skip_f32, float* out_f32, void* stream) {
  if (epi != EPI_BIAS_RELU && epi != EPI_BIAS_SKIP) return DFIR_ERR_ARG;
  if (epi == EPI_BIAS_SKIP && skip_f32 == nullptr) return DFIR_ERR_ARG;
  if (out_bf16 == nullptr || r_bf16 == nullptr || x_in == nullptr) return DFIR_ERR_ARG;
  ConvTcDesc d{};
  d.B = B; d.H = H; d.W = W; d.cin_total = 64; d.cin_off = 0; d.cout = 64; d.epi = epi; d.in_mode = IN_FUSED;
  d.r_bf16 = r_bf16; d.xin_f32 = x_in; d.xout_f32 = x_out; d.pool_rows = const_cast<float*>(pool_rows);
  d.ca_style = style; d.ca_params = ca_params; d.ca_R = R; d.ca_M = M; d.ca_A = A; d.attributes = attributes;
  d.sq = sq; d.res_scale = res_scale;
  d.wpacked = wpacked; d.bias = bias; d.out_bf16 = out_bf16;
  d.out_pix_stride = 128; d.out_row_stride = static_cast<long long>(W) * 128;
  d.out_img_stride = static_cast<long long>(H) * W * 128;
  d.skip_f32 = skip_f32; d.out_f32 = out_f32;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return DFIR_ERR_CUDA;
  d.num_sms = sms;
  return conv3x3_c64_tc(d, S(stream));
}

int dfir_conv3x3_c64_stats(const void* in_bf16, const void* wpacked, const float* bias, int B, int H, int W,
                           void* out_bf16, float* pool_rows, float* col_first, float* col_last, void* stream) {
  if (in_bf16 == nullptr || out_bf16 == nullptr) return DFIR_ERR_ARG;
  ConvTcDesc d{};
  d.B = B; d.H = H; d.W = W; d.cin_total = 64; d.cin_off = 0; d.cout = 64; d.epi = EPI_RELU_STATS; d.in_mode = IN_TMA;
  d.in_bf16 = in_bf16; d.wpacked = wpacked; d.bias = bias; d.out_bf16 = out_bf16;
  d.out_pix_stride = 128; d.out_row_stride = static_cast<long long>(W) * 128;
  d.out_img_stride = static_cast<long long>(H) * W * 128;
  d.pool_rows = pool_rows; d.col_first = col_first; d.col_last = col_last;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return DFIR_ERR_CUDA;
  d.num_sms = sms;
  return conv3x3_c64_tc(d, S(stream));
}

int dfir_conv3x3_c64_stats_fx(const void* in_bf16, const void* wpacked, const float* bias, int B, int H, int W,
                              void* out_bf16, long long* istats, int warp_autonomous, void* stream) {
  if (in_bf16 == nullptr || out_bf16 == nullptr || istats == nullptr) return DFIR_ERR_ARG;
  ConvTcDesc d{};
  d.B = B; d.H = H; d.W = W; d.cin_total = 64; d.cin_off = 0; d.cout = 64; d.in_mode = IN_TMA;
  d.epi = warp_autonomous ? EPI_RELU_STATS_W : EPI_RELU_STATS;
  d.in_bf16 = in_bf16; d.wpacked = wpacked; d.bias = bias; d.out_bf16 = out_bf16;
  d.out_pix_stride = 128; d.out_row_stride = static_cast<long long>(W) * 128;
  d.out_img_stride = static_cast<long long>(H) * W * 128;
  d.istats = istats;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return DFIR_ERR_CUDA;
  d.num_sms = sms;
  return conv3x3_c64_tc(d, S(stream));
}

int dfir_ca_from_stats(const float* pool_rows, const float* col_first, const float* col_last, const void* w2_packed,
                       const float* bias2, int style, const float* ca_params, int R, int M, int A,
                       const float* attributes, const float* sq, float* svec, int B, int H, int W, void* stream) {
  AttnParams ap{};
  ap.style = style; ap.C = 64; ap.R = R; ap.M = M; ap.A = A; ap.w[0] = ca_params;
  if (pool_rows == nullptr || col_first == nullptr || col_last == nullptr || w2_packed == nullptr ||
      ca_params == nullptr || svec == nullptr)
    return DFIR_ERR_ARG;
  return ca_from_stats(pool_rows, col_first, col_last, w2_packed, bias2, ap, attributes, sq, svec, B, H, W, S(stream));
}

int dfir_conv3x3_c64_scale_skip(const void* in_bf16, const void* wpacked, const float* bias, int B, int H, int W,
                                const float* svec, const float* skip_f32, float* out_f32, void* out_bf16,
                                const float* pool_rows, const float* col_first, const float* col_last, int style,
                                const float* ca_params, int R, int M, int A, const float* attributes,
                                const float* sq, void* stream) {
  if (in_bf16 == nullptr || out_bf16 == nullptr || skip_f32 == nullptr) return DFIR_ERR_ARG;
  ConvTcDesc d{};
  d.B = B; d.H = H; d.W = W; d.cin_total = 64; d.cin_off = 0; d.cout = 64; d.epi = EPI_SCALE_SKIP; d.in_mode = IN_TMA;
  d.in_bf16 = in_bf16; d.wpacked = wpacked; d.bias = bias; d.out_bf16 = out_bf16;
  d.out_pix_stride = 128; d.out_row_stride = static_cast<long long>(W) * 128;
  d.out_img_stride = static_cast<long long>(H) * W * 128;
  d.svec = svec; d.skip_f32 = skip_f32; d.out_f32 = out_f32;
  d.pool_rows = const_cast<float*>(pool_rows); d.col_first = const_cast<float*>(col_first);
  d.col_last = const_cast<float*>(col_last);
  d.ca_style = style; d.ca_params = ca_params; d.ca_R = R; d.ca_M = M; d.ca_A = A; d.attributes = attributes; d.sq = sq;
  d.epi_stats = style != DFIR_STYLE_NONE ? 1 : 0;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return DFIR_ERR_CUDA;
  d.num_sms = sms;
  return conv3x3_c64_tc(d, S(stream));
}

static int scale_skip_hl_any(int epi, const void* in_bf16, const void* wpacked, const float* bias, int B, int H, int W,
                                   const float* svec, const void* skip_hi, const void* skip_lo, void* out_hi, void* out_lo,
                                   const float* pool_rows, const float* col_first, const float* col_last, int style,
                                   const float* ca_params, int R, int M, int A, const float* attributes,
                                   const float* sq, int descending, void* stream) {
  if (in_bf16 == nullptr || out_hi == nullptr || skip_hi == nullptr || skip_lo == nullptr) return DFIR_ERR_ARG;
  ConvTcDesc d{};
  d.B = B; d.H = H; d.W = W; d.cin_total = 64; d.cin_off = 0; d.cout = 64; d.epi = epi; d.in_mode = IN_TMA;
  d.in_bf16 = in_bf16; d.wpacked = wpacked; d.bias = bias; d.out_bf16 = out_hi; d.out_lo = out_lo;
  d.skip_hi = skip_hi; d.skip_lo = skip_lo; d.flip = descending ? 1 : 0;
  d.out_pix_stride = 128; d.out_row_stride = static_cast<long long>(W) * 128;
  d.out_img_stride = static_cast<long long>(H) * W * 128;
  d.svec = svec;
  d.pool_rows = const_cast<float*>(pool_rows); d.col_first = const_cast<float*>(col_first);
  d.col_last = const_cast<float*>(col_last);
  d.ca_style = style; d.ca_params = ca_params; d.ca_R = R; d.ca_M = M; d.ca_A = A; d.attributes = attributes; d.sq = sq;
  d.epi_stats = style != DFIR_STYLE_NONE ? 1 : 0;
  // statistics as the 64-bit fixed-point image sums of dfir_conv3x3_c64_stats_fx (what the network schedule uses):
  // col_first == col_last == NULL and `pool_rows` points at the [B][9][64] int64 sums
  if (d.epi_stats && pool_rows != nullptr && col_first == nullptr && col_last == nullptr) {
    d.epi_stats = 2;
    d.istats = reinterpret_cast<long long*>(const_cast<float*>(pool_rows));
    d.pool_rows = nullptr;
  }
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return DFIR_ERR_CUDA;
  d.num_sms = sms;
  return conv3x3_c64_tc(d, S(stream));
}

int dfir_conv3x3_c64_scale_skip_hl(const void* in_bf16, const void* wpacked, const float* bias, int B, int H, int W,
                                   const float* svec, const void* skip_hi, const void* skip_lo, void* out_hi, void* out_lo,
                                   const float* pool_rows, const float* col_first, const float* col_last, int style,
                                   const float* ca_params, int R, int M, int A, const float* attributes,
                                   const float* sq, int descending, void* stream) {
  return scale_skip_hl_any(EPI_SCALE_SKIP_HL, in_bf16, wpacked, bias, B, H, W, svec, skip_hi, skip_lo, out_hi, out_lo, pool_rows,
                           col_first, col_last, style, ca_params, R, M, A, attributes, sq, descending, stream);
}

int dfir_conv3x3_c64_accumulate_hl8(const void* in_bf16, const void* wpacked, const float* bias, int B, int H, int W,
                                    const float* svec, const void* skip_hi, const void* skip_lo8, void* out_hi, void* out_lo8,
                                    int relu, void* stream) {
  if (in_bf16 == nullptr || out_hi == nullptr || skip_hi == nullptr || skip_lo8 == nullptr) return DFIR_ERR_ARG;
  ConvTcDesc d{};
  d.B = B; d.H = H; d.W = W; d.cin_total = 64; d.cin_off = 0; d.cout = 64; d.epi = EPI_SCALE_SKIP_HL8; d.in_mode = IN_TMA;
  d.in_bf16 = in_bf16; d.wpacked = wpacked; d.bias = bias; d.out_bf16 = out_hi; d.out_lo = out_lo8;
  d.skip_hi = skip_hi; d.skip_lo = skip_lo8; d.relu_out = relu ? 1 : 0;
  d.out_pix_stride = 128; d.out_row_stride = static_cast<long long>(W) * 128;
  d.out_img_stride = static_cast<long long>(H) * W * 128;
  d.svec = svec;
  DFIR_TRY(dfir_check_device());
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return DFIR_ERR_CUDA;
  d.num_sms = sms;
  return conv3x3_c64_tc(d, S(stream));
}

int dfir_stream_encode_hl8(const float* x, void* hi, void* lo8, long long n, void* stream) {
  if (x == nullptr || hi == nullptr || lo8 == nullptr) return DFIR_ERR_ARG;
  return stream_encode_hl8(x, hi, lo8, n, S(stream));
}

int dfir_stream_decode_hl8(const void* hi, const void* lo8, float* x, long long n, void* stream) {
  if (x == nullptr || hi == nullptr || lo8 == nullptr) return DFIR_ERR_ARG;
  return stream_decode_hl8(hi, lo8, x, n, S(stream));
}

int dfir_conv3x3_c64_scale_skip_hl8(const void* in_bf16, const void* wpacked, const float* bias, int B, int H, int W,
                                    const float* svec, const void* skip_hi, const void* skip_lo8, void* out_hi, void* out_lo8,
                                    const float* pool_rows, const float* col_first, const float* col_last, int style,
                                    const float* ca_params, int R, int M, int A, const float* attributes,
                                    const float* sq, int descending, void* stream) {
  return scale_skip_hl_any(EPI_SCALE_SKIP_HL8, in_bf16, wpacked, bias, B, H, W, svec, skip_hi, skip_lo8, out_hi, out_lo8,
                           pool_rows, col_first, col_last, style, ca_params, R, M, A, attributes, sq, descending, stream);
}

int dfir_conv3x3_c64_accumulate(const void* in_bf16, const void* wpacked, const float* bias, int B, int H, int W,
                                const float* svec, const float* skip_f32, float* out_f32, void* out_bf16, int relu,
                                void* stream) {
  if (in_bf16 == nullptr || wpacked == nullptr || out_bf16 == nullptr) return DFIR_ERR_ARG;
  ConvTcDesc d{};
  d.B = B; d.H = H; d.W = W; d.cin_total = 64; d.cin_off = 0; d.cout = 64; d.epi = EPI_SCALE_SKIP; d.in_mode = IN_TMA;
  d.in_bf16 = in_bf16; d.wpacked = wpacked; d.bias = bias; d.out_bf16 = out_bf16;
  d.out_pix_stride = 128; d.out_row_stride = static_cast<long long>(W) * 128;
  d.out_img_stride = static_cast<long long>(H) * W * 128;
  d.svec = svec; d.skip_f32 = skip_f32; d.out_f32 = out_f32; d.relu_out = relu ? 1 : 0;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return DFIR_ERR_CUDA;
  d.num_sms = sms;
  return conv3x3_c64_tc(d, S(stream));
}

int dfir_conv3x3_c64_tail(const void* in_bf16, const void* wpacked, const float* bias, int B, int H, int W, int cout,
                          float* out_nchw, int accumulate, void* stream) {
  if (in_bf16 == nullptr || wpacked == nullptr || out_nchw == nullptr || cout < 1 || cout > 16) return DFIR_ERR_ARG;
  ConvTcDesc d{};
  d.B = B; d.H = H; d.W = W; d.cin_total = 64; d.cin_off = 0; d.cout = cout; d.epi = EPI_TAIL_NCHW; d.in_mode = IN_TMA;
  d.in_bf16 = in_bf16; d.wpacked = wpacked; d.bias = bias; d.out_f32 = out_nchw; d.tail_accumulate = accumulate ? 1 : 0;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return DFIR_ERR_CUDA;
  d.num_sms = sms;
  return conv3x3_c64_tc(d, S(stream));
}

int dfir_conv3x3_f32(const float* in, const float* w_packed, const float* bias, const float* skip, float* out, int B,
                     int H, int W, int Cin, int Cout, int relu, int ps_r, int out_nchw, void* stream) {
  return conv3x3_f32(in, w_packed, bias, skip, out, B, H, W, Cin, Cout, relu, ps_r, out_nchw, S(stream));
}

int dfir_head_conv(const float* x_nchw, const float* w_packed, const float* bias, float* out_f32, void* out_bf16,
                   int B, int Cin, int H, int W, int Cout, void* stream) {
  return head_conv(x_nchw, w_packed, bias, out_f32, reinterpret_cast<__nv_bfloat16*>(out_bf16), B, Cin, H, W, Cout,
                   S(stream));
}

int dfir_meta_attention(const float* meta, const float* w1, const float* b1, const float* w2, const float* b2,
                        float* out, int nblk, int B, int M, int Hid, int C, int relu, const int* blk_enabled,
                        float out_scale, void* stream) {
  return meta_attention(meta, w1, b1, w2, b2, out, nblk, B, M, Hid, C, relu, blk_enabled, out_scale, S(stream));
}

int dfir_ca_scale_residual(const void* r, int r_is_bf16, const float* x_in, const float* pool_rows, int pool_nrows,
                           int style, const float* ca_params, int C, int R, int M, int A, const float* attributes,
                           const float* sq, float res_scale, float* x_out, void* x_out_bf16, int B, int H, int W,
                           void* stream) {
  AttnParams ap{};
  ap.style = style; ap.C = C; ap.R = R; ap.M = M; ap.A = A; ap.w[0] = ca_params;
  if (style != DFIR_STYLE_NONE && (pool_rows == nullptr || ca_params == nullptr)) return DFIR_ERR_ARG;
  return scale_residual(r, r_is_bf16, x_in, pool_rows, pool_nrows, ap, attributes, sq, res_scale, x_out,
                        reinterpret_cast<__nv_bfloat16*>(x_out_bf16), B, H, W, C, S(stream));
}

int dfir_ca_pa_scale_residual(const void* r, int r_is_bf16, const float* x_in, const float* pool_rows, int pool_nrows,
                              int style, const float* ca_params, const float* pa_params, int R, int M, int A,
                              const float* attributes, const float* sq, float* x_out, void* x_out_bf16, int B, int H,
                              int W, void* stream) {
  AttnParams ap{};
  ap.style = style; ap.C = 64; ap.R = R; ap.M = M; ap.A = A; ap.w[0] = ca_params;
  if (style == DFIR_STYLE_NONE || pool_rows == nullptr || ca_params == nullptr || pa_params == nullptr) return DFIR_ERR_ARG;
  return scale_residual(r, r_is_bf16, x_in, pool_rows, pool_nrows, ap, attributes, sq, 1.f, x_out,
                        reinterpret_cast<__nv_bfloat16*>(x_out_bf16), B, H, W, 64, S(stream), nullptr, pa_params);
}

int dfir_postprocess_rgb(const float* x_nchw, float* rgb_clipped, float* ycbcr, int B, long long HW, float lo, float hi,
                         void* stream) {
  if (x_nchw == nullptr || rgb_clipped == nullptr || ycbcr == nullptr) return DFIR_ERR_ARG;
  return postprocess_rgb(x_nchw, rgb_clipped, ycbcr, B, HW, lo, hi, S(stream));
}

size_t dfir_postprocess_u8_scratch_bytes(int B, long long HW) {
  return B <= 0 || HW <= 0 ? 0 : static_cast<size_t>(B) * postprocess_u8_blocks(B, HW) * sizeof(double);
}

int dfir_postprocess_u8(const float* x_nchw, const float* hr_nchw, unsigned char* rgb_u8, unsigned char* ycbcr_u8,
                        float* y_psnr, void* scratch, size_t scratch_bytes, int B, long long HW, void* stream) {
  if (x_nchw == nullptr || (rgb_u8 == nullptr && ycbcr_u8 == nullptr && hr_nchw == nullptr)) return DFIR_ERR_ARG;
  if (hr_nchw != nullptr && scratch_bytes < dfir_postprocess_u8_scratch_bytes(B, HW)) return DFIR_ERR_WORKSPACE;
  return postprocess_u8(x_nchw, hr_nchw, rgb_u8, ycbcr_u8, y_psnr, reinterpret_cast<double*>(scratch), B, HW, S(stream));
}

int dfir_pool_rows_f32(const float* in, float* pool_rows, int B, int H, int W, int C, void* stream) {
  return pool_rows_f32(in, pool_rows, B, H, W, C, S(stream));
}

size_t dfir_qrcan_workspace_bytes(const dfir_qrcan_net* net, int B, int H, int W, int precision) {
  if (net == nullptr || B <= 0 || H <= 0 || W <= 0) return 0;
  const int Bc = net->chunk_images > 0 ? std::min(net->chunk_images, B) : auto_chunk(B, H, W, precision);
  const int ways = split_ways(net, Bc, precision);
  const int Bs = (Bc + ways - 1) / ways;
  return ways * align256(carve_qrcan(net, B, Bs, H, W, precision, nullptr).total);
}

long long dfir_qrcan_launch_count(const dfir_qrcan_net* net, int B, int H, int W, int precision) {
  if (net == nullptr || B <= 0) return 0;
  const int Bc = net->chunk_images > 0 ? std::min(net->chunk_images, B) : auto_chunk(B, H, W, precision);
  const long long chunks = (B + Bc - 1) / Bc;
  int r = 0;
  const int nup = up_stages(net->scale, &r);
  const long long nb = net->n_blocks, ng = net->n_groups;
  long long per_chunk;
  if (precision == DFIR_PREC_BF16_TC) {
    per_chunk = 1 + ng * (nb * ((net->schedule == 2 || (net->schedule == 3 && net->style != DFIR_STYLE_NONE) || net->pa_blob != nullptr) ? 3 : 2) + (net->no_group_conv ? 0 : 1)) + 1 +
                static_cast<long long>(nup) * r * r + 1;
  } else {
    const long long pool = net->style != DFIR_STYLE_NONE ? 1 : 0;
    per_chunk = 1 + ng * (nb * (3 + pool) + (net->no_group_conv ? 0 : 2)) + 1 + nup + 1;  // group tail = conv + copy
  }
  return chunks * per_chunk + (net->any_q ? 1 : 0);
}

int dfir_qrcan_stages(const dfir_qrcan_net* net, int stages, int g_begin, int g_end, const float* x_nchw,
                      const float* attributes, const float* feat_in_f32, float* group_out_f32, float* feat_out_f32,
                      float* out_nchw, int B, int H, int W, int precision, void* workspace, size_t workspace_bytes,
                      void* stream) {
  if (net == nullptr || B <= 0 || H <= 0 || W <= 0 || (stages & ~ST_ALL) != 0 || stages == 0) return DFIR_ERR_ARG;
  if ((stages & ST_HEAD) && x_nchw == nullptr) return DFIR_ERR_ARG;
  if ((stages & ST_GROUPS) && (g_begin < 0 || g_end > net->n_groups || g_begin > g_end || attributes == nullptr))
    return DFIR_ERR_ARG;
  if ((stages & ST_UPSAMPLE) && out_nchw == nullptr) return DFIR_ERR_ARG;
  int r = 0;
  if (up_stages(net->scale, &r) < 0) return DFIR_ERR_ARG;
  if (precision == DFIR_PREC_BF16_TC && net->n_feats != 64) return DFIR_ERR_ARG;
  if (precision != DFIR_PREC_BF16_TC && precision != DFIR_PREC_FP32_SIMT) return DFIR_ERR_ARG;
  DFIR_TRY(dfir_check_device());
  const int Bc = net->chunk_images > 0 ? std::min(net->chunk_images, B) : auto_chunk(B, H, W, precision);
  if (Bc != B) return DFIR_ERR_ARG;  // staged execution keeps its state in the workspace: one pass only
  QrcanWs w = carve_qrcan(net, B, Bc, H, W, precision, workspace);
  if (workspace == nullptr || w.total > workspace_bytes) return DFIR_ERR_WORKSPACE;
  cudaStream_t st = S(stream);
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return DFIR_ERR_CUDA;
  const int C = net->n_feats;
  const size_t feat_n = static_cast<size_t>(B) * H * W * C;
  const bool tc = precision == DFIR_PREC_BF16_TC;
  StageArgs sa;
  sa.stages = stages; sa.g_begin = g_begin; sa.g_end = g_end; sa.group_out = group_out_f32;
  if (feat_in_f32 != nullptr) {
    if (stages & ST_GROUPS) {  // external features become the stream entering g_begin
      if (cudaMemcpyAsync(w.XA, feat_in_f32, feat_n * 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess) return DFIR_ERR_CUDA;
      if (tc) DFIR_TRY(f32_to_bf16(w.XA, w.XAbf, static_cast<long long>(feat_n), st));
      sa.from_xa = 1;
    } else if (stages & ST_TRUNK_TAIL) {
      if (cudaMemcpyAsync(w.XA, feat_in_f32, feat_n * 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess) return DFIR_ERR_CUDA;
      if (tc) DFIR_TRY(f32_to_bf16(w.XA, w.XAbf, static_cast<long long>(feat_n), st));
    } else if (stages & ST_UPSAMPLE) {  // external features are the trunk output feeding the upsampler
      if (tc) DFIR_TRY(f32_to_bf16(feat_in_f32, w.XBbf, static_cast<long long>(feat_n), st));
      else if (cudaMemcpyAsync(w.R32, feat_in_f32, feat_n * 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess) return DFIR_ERR_CUDA;
    }
  }
  if ((stages & ST_GROUPS) && net->any_q) {
    DFIR_TRY(meta_attention(attributes, net->meta_w1, net->meta_b1, net->meta_w2, net->meta_b2, w.sq,
                            net->n_groups * net->n_blocks, B, net->num_metadata, net->meta_hidden, net->n_feats,
                            net->meta_relu, net->q_enabled, net->style == DFIR_STYLE_NONE ? net->res_scale : 1.f, st));
  }
  if (tc) DFIR_TRY(qrcan_forward_bf16(net, x_nchw, attributes, out_nchw, B, B, 0, H, W, w, sms, sa, st));
  else DFIR_TRY(qrcan_forward_f32(net, x_nchw, attributes, out_nchw, B, B, 0, H, W, w, sa, st));
  if (feat_out_f32 != nullptr) {
    const float* src = nullptr;
    if ((stages & ST_GROUPS) && g_end > g_begin) src = (net->no_group_conv && net->n_blocks > 0) ? w.XB : w.XA;
    else if (stages & ST_HEAD) src = w.Hh;
    if (src == nullptr) return DFIR_ERR_ARG;
    if (cudaMemcpyAsync(feat_out_f32, src, feat_n * 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess) return DFIR_ERR_CUDA;
  }
  return DFIR_OK;
}

int dfir_channel_scale(const float* x, const float* svec, const float* add, float alpha, float* out, int B, int HW,
                       int C, void* stream) {
  return channel_scale(x, svec, add, alpha, out, B, HW, C, S(stream));
}

size_t dfir_lam_scratch_bytes(int B, int N) { return lam_scratch_floats(B, N) * 4; }
int dfir_lam(const float* stack, long long map_stride_elems, float gamma, float* out, void* scratch, int N, int B,
             int HW, int C, void* stream) {
  return lam_forward(stack, map_stride_elems, gamma, out, reinterpret_cast<float*>(scratch), N, B, HW, C, S(stream));
}

int dfir_csam(const float* x, const float* w27, float bias, float gamma, float* out, int B, int H, int W, int C,
              void* stream) {
  return csam_forward(x, w27, bias, gamma, out, B, H, W, C, S(stream));
}

size_t dfir_soca_scratch_bytes(int B) { return soca_scratch_floats(B) * 4; }
int dfir_soca(const float* x, const float* mlp_params, int R, float* svec, void* scratch, int B, int H, int W, int C,
              void* stream) {
  return soca_forward(x, mlp_params, R, svec, reinterpret_cast<float*>(scratch), B, H, W, C, S(stream));
}

int dfir_batch_blur(const float* x_nchw, const float* kernels, int kernel_per_image, const float* noise,
                    const float* noise_sigma, float* out_nchw, int B, int C, int H, int W, int l, int clamp01, void* stream) {
  if (x_nchw == nullptr || kernels == nullptr || out_nchw == nullptr || x_nchw == out_nchw) return DFIR_ERR_ARG;
  if ((noise == nullptr) != (noise_sigma == nullptr)) return DFIR_ERR_ARG;
  return batch_blur(x_nchw, kernels, kernel_per_image, noise, noise_sigma, out_nchw, B, C, H, W, l, clamp01, S(stream));
}
int dfir_pca_encode(const float* kernels, const float* pca_matrix, const float* noise_sigma, float* code, int B, int l,
                    int k, void* stream) {
  if (kernels == nullptr || pca_matrix == nullptr || code == nullptr) return DFIR_ERR_ARG;
  return pca_encode(kernels, pca_matrix, noise_sigma, code, B, l * l, k, S(stream));
}

size_t dfir_covpool_scratch_bytes(int B) { return covpool_scratch_floats(B) * 4; }
int dfir_covpool(const float* x, float* cov, void* scratch, size_t scratch_bytes, int B, int H, int W, int C, int crop1000,
                 void* stream) {
  if (x == nullptr || cov == nullptr || B < 0 || H <= 0 || W <= 0) return DFIR_ERR_ARG;
  if (scratch == nullptr || scratch_bytes < dfir_covpool_scratch_bytes(B)) return DFIR_ERR_WORKSPACE;
  return covpool_forward(x, cov, reinterpret_cast<float*>(scratch), B, H, W, C, crop1000, S(stream));
}
int dfir_covpool_backward(const float* x, const float* grad_cov, float* grad_x, void* scratch, size_t scratch_bytes, int B,
                          int H, int W, int C, int crop1000, void* stream) {
  if (x == nullptr || grad_cov == nullptr || grad_x == nullptr || B < 0 || H <= 0 || W <= 0) return DFIR_ERR_ARG;
  if (scratch == nullptr || scratch_bytes < dfir_covpool_scratch_bytes(B)) return DFIR_ERR_WORKSPACE;
  return covpool_backward(x, grad_cov, grad_x, reinterpret_cast<float*>(scratch), B, H, W, C, crop1000, S(stream));
}
size_t dfir_sqrtm_scratch_bytes(int B, int iters) { return sqrtm_scratch_floats(B, iters) * 4; }
int dfir_sqrtm(const float* cov, float* out, int B, int C, int iters, void* stream) {
  if (cov == nullptr || out == nullptr || B < 0) return DFIR_ERR_ARG;
  return sqrtm_forward(cov, out, B, C, iters, S(stream));
}
int dfir_sqrtm_backward(const float* cov, const float* grad_out, float* grad_in, void* scratch, size_t scratch_bytes, int B,
                        int C, int iters, void* stream) {
  if (cov == nullptr || grad_out == nullptr || grad_in == nullptr || B < 0) return DFIR_ERR_ARG;
  if (scratch == nullptr || scratch_bytes < dfir_sqrtm_scratch_bytes(B, iters)) return DFIR_ERR_WORKSPACE;
  return sqrtm_backward(cov, grad_out, grad_in, reinterpret_cast<float*>(scratch), B, C, iters, S(stream));
}

size_t dfir_nonlocal_scratch_bytes(int B, int H, int W) { return nonlocal_scratch_floats(B, H, W) * 4; }
int dfir_nonlocal(const float* x, const float* w_tpg, const float* b_tpg, const float* w_out, const float* b_out,
                  float* out, void* scratch, int B, int H, int W, int C, void* stream) {
  return nonlocal_forward(x, w_tpg, b_tpg, w_out, b_out, out, reinterpret_cast<float*>(scratch), B, H, W, C, S(stream));
}

int dfir_qrcan_forward(const dfir_qrcan_net* net, const float* x_nchw, const float* attributes, float* out_nchw, int B,
                       int H, int W, int precision, void* workspace, size_t workspace_bytes, void* stream) {
  if (net == nullptr || x_nchw == nullptr || out_nchw == nullptr || attributes == nullptr) return DFIR_ERR_ARG;
  if (B <= 0 || H <= 0 || W <= 0) return DFIR_ERR_ARG;
  int r = 0;
  if (up_stages(net->scale, &r) < 0) return DFIR_ERR_ARG;
  if (precision == DFIR_PREC_BF16_TC && net->n_feats != 64) return DFIR_ERR_ARG;
  if (net->n_feats % 8 != 0 || net->n_feats > 256 || 256 % net->n_feats != 0) return DFIR_ERR_ARG;
  if (precision != DFIR_PREC_BF16_TC && precision != DFIR_PREC_FP32_SIMT) return DFIR_ERR_ARG;
  DFIR_TRY(dfir_check_device());
  const int Bc = net->chunk_images > 0 ? std::min(net->chunk_images, B) : auto_chunk(B, H, W, precision);
  const int ways = split_ways(net, Bc, precision);
  const int Bs = (Bc + ways - 1) / ways;  // images per concurrent sub-pass
  QrcanWs w = carve_qrcan(net, B, Bs, H, W, precision, workspace);
  const size_t sub_bytes = align256(w.total);
  if (workspace == nullptr || ways * sub_bytes > workspace_bytes) return DFIR_ERR_WORKSPACE;
  cudaStream_t st = S(stream);
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return DFIR_ERR_CUDA;
  if (net->any_q) {
    DFIR_TRY(meta_attention(attributes, net->meta_w1, net->meta_b1, net->meta_w2, net->meta_b2, w.sq,
                            net->n_groups * net->n_blocks, B, net->num_metadata, net->meta_hidden, net->n_feats,
                            net->meta_relu, net->q_enabled, net->style == DFIR_STYLE_NONE ? net->res_scale : 1.f, st));
  }
  StageArgs all;
  all.g_end = net->n_groups;
  if (ways > 1) {
    SplitCtx* sc = split_ctx(dev);
    if (sc == nullptr) return DFIR_ERR_CUDA;
    QrcanWs ws[SplitCtx::kMaxWays];
    for (int i = 0; i < ways; ++i) {
      ws[i] = carve_qrcan(net, B, Bs, H, W, precision, reinterpret_cast<uint8_t*>(workspace) + i * sub_bytes);
      ws[i].sq = w.sq;  // the meta-attention vectors of the whole batch, written once above
    }
    const int sub_sms = std::max(1, sms / ways);
    for (int b0 = 0; b0 < B; b0 += Bc) {
      const int bc = std::min(Bc, B - b0);
      if (cudaEventRecord(sc->fork, st) != cudaSuccess) return DFIR_ERR_CUDA;
      int used = 0;
      for (int i = 0; i < ways; ++i) {
        const int s0 = i * Bs, sn = std::min(Bs, bc - s0);
        if (sn <= 0) break;
        cudaStream_t si = i == 0 ? st : sc->side[i - 1];
        if (i > 0 && cudaStreamWaitEvent(si, sc->fork, 0) != cudaSuccess) return DFIR_ERR_CUDA;
        StageArgs sa = all;
        sa.phase_wait = i > 0 ? sc->phase[i - 1] : nullptr;
        sa.phase_record = (i + 1 < ways && (i + 1) * Bs < bc) ? sc->phase[i] : nullptr;
        DFIR_TRY(qrcan_forward_bf16(net, x_nchw, attributes, out_nchw, B, sn, b0 + s0, H, W, ws[i], sub_sms, sa, si));
        if (i > 0 && cudaEventRecord(sc->join[i - 1], si) != cudaSuccess) return DFIR_ERR_CUDA;
        used = i + 1;
      }
      for (int i = 1; i < used; ++i)
        if (cudaStreamWaitEvent(st, sc->join[i - 1], 0) != cudaSuccess) return DFIR_ERR_CUDA;
    }
    return DFIR_OK;
  }
  for (int b0 = 0; b0 < B; b0 += Bc) {
    const int bc = std::min(Bc, B - b0);
    if (precision == DFIR_PREC_BF16_TC)
      DFIR_TRY(qrcan_forward_bf16(net, x_nchw, attributes, out_nchw, B, bc, b0, H, W, w, sms, all, st));
    else
      DFIR_TRY(qrcan_forward_f32(net, x_nchw, attributes, out_nchw, B, bc, b0, H, W, w, all, st));
  }
  return DFIR_OK;
}

}  // extern "C"
