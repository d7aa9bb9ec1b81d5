// 3x3 convolution, 64 input channels, as an implicit GEMM on the sm_100a tensor cores.
//
//   reference op : default_conv = nn.Conv2d(k=3, stride 1, zero pad 1, bias)
//                  (/root/reference/Code/SISR/models/advanced/common.py:5-8)
//
// Data layout (HBM): activations NHWC bf16 (one pixel = 64 ch = 128 B = exactly one 128B-swizzle row),
// weights pre-packed per tap as [tap][cout][cin] bf16 in the UMMA K-major SWIZZLE_128B byte order.
//
// GEMM view: D[M = 128 pixels of one image row][N = cout] += A[M][K = 64 cin] * B[N][K] for each of the 9
// taps (K total = 576).  No im2col is ever materialised.  A persistent CTA streams input rows into a small
// shared-memory ring (TMA: one box = one row segment of 130 pixels incl. the x halo; out-of-bounds pixels
// and rows are zero-filled by the TMA unit, which IS the conv's zero padding).
//
// Operand placement (measured, profiles/r01_*): the tensor core fetches shared-memory operands at only
// ~64 B/clk/SM, so an SS-mode MMA of this shape (A 4 KB + B 2 KB per 32-clk instruction) is operand-fetch
// bound at 1/3 of peak.  The A operand therefore lives in TENSOR MEMORY: four loader warps copy every ring
// row into TMEM three times, shifted by dx = 0,1,2 pixels (lane m <- pixel m + dx; 32 columns = the pixel's
// 64 bf16 channels), with ordinary swizzle-aware LDS.128 + tcgen05.st.  A tap's A operand is then just a
// TMEM address: dy picks the row slot, dx the copy, k the 8-column slice; the MMAs (tcgen05.mma, A from
// TMEM) only pull the 2 KB weight tile from shared memory.  The weights of the layer (9 x cout x 128 B)
// stay resident in shared memory for the CTA's whole life.  TMEM budget (512 columns): 2 accumulators
// (2 x 64) + 4 row slots x 3 copies x 32 columns.
//
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + tcgen05.mma issuer, warps 2-5 = epilogue
// (TMEM -> registers -> fused bias/ReLU/skip/pool -> swizzled smem -> TMA store), warps 6-9 = A loaders.
// EPI_SCALE_SKIP (the fp32 residual-stream epilogue: scale, skip add, fp32 + bf16 stores; also the fp32-accumulating
// data-gradient conv of the backward pass) adds a second epilogue group, warps 10-13: the two groups drain alternate
// accumulators, each in two 32-channel halves through a 16 KB fp32 transposition tile; both halves of the fp32 skip row
// have their own 16 KB cp.async buffer per group, so a half is requested a full pass before it is read.  The inference
// case (full row, skip + fp32 output, nothing else) runs a straight-line specialisation of the pass: the loop is
// instruction-fetch sensitive.  EPI_RELU_STATS (conv1 + the statistics of pool-by-linearity) uses the same two groups:
// its 62 shuffles per thread and row do not fit one group's share of the MMA pace.
//
// Environment switches (A/B measurements): DFIR_PDL, DFIR_L2_POLICY (L2 eviction-priority hints, default
// none), DFIR_WPREFETCH (next layer's weights into L2, default off), DFIR_NUM_SMS; DFIR_DEBUG_PROBE timing experiments
// exist only in a `make PROBES=1` build (bit 16384, fast path off, always).
//
// The kernel is launched with the programmatic-dependent-launch attribute: after its set-up it lets the next kernel of
// the stream become resident as CTAs retire (griddepcontrol.launch_dependents) and waits for its own predecessor
// (griddepcontrol.wait) before touching anything that predecessor wrote; the weight load precedes the wait.
//
// Backward-pass use (csrc/train_api.cu): the data gradient of a conv is this kernel on the transposed, 180-degree
// rotated weights — with EPI_RELU_MASK (times the saved ReLU mask) or EPI_SCALE_SKIP (fp32 accumulate); strided TMA
// input views address one PixelShuffle phase of an output gradient.
//
// Input modes
//   IN_TMA   : the input rows are bf16 NHWC in HBM/L2 and arrive by TMA (320 threads).
//   IN_FUSED : the input of the conv is x' = r * s + x, i.e. the previous block's channel-attention scale
//              and residual add (QRCAB: `res * y` twice and `res += x`, architectures.py:127,172-180;
//              q_layer.py:43).  Eight extra warps (two groups that alternate rows) read r (bf16) and x
//              (fp32 residual stream), form x', write it back as the new fp32 stream for the rows the CTA
//              owns, and deposit bf16(x') into the swizzled ring slot.  The separate elementwise pass (and
//              its 12 B/element of traffic) disappears (576 threads).
//              The attention vector s = CA_style(pooled mean, attributes) * meta_scale (attn.cuh) is
//              evaluated by the same warps in the kernel prologue, hidden behind the weight load.
#include "ptx.cuh"
#include "kernels.h"
#include "attn.cuh"

#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>

namespace dfir {

__device__ unsigned int g_dfir_progress[16];
// `make PROBES=1` + DFIR_DEBUG_PROBE bit 32768: clock64 time stamps of one CTA's pipeline events, [16 kinds][64 rows]
// (dfir_debug_trace reads them back; tools/trace_conv.py prints who waits for whom).
__device__ unsigned long long g_dfir_trace[16 * 64];
#ifdef DFIR_PROBES
#define DFIR_TRACE(kind, idx)                                                                   \
  do {                                                                                          \
    if (trace_on && lane == 0 && (idx) >= 0 && (idx) < 64) g_dfir_trace[(kind) * 64 + (idx)] = clock64(); \
  } while (0)
#define DFIR_TRACE1(kind, idx)                                                                  \
  do {                                                                                          \
    if (trace_on && (idx) >= 0 && (idx) < 64) g_dfir_trace[(kind) * 64 + (idx)] = clock64();   \
  } while (0)
#else
#define DFIR_TRACE(kind, idx) do { } while (0)
#define DFIR_TRACE1(kind, idx) do { } while (0)
#endif

using namespace ptx;

namespace {

constexpr int kSlots = 3;                    // smem input-row ring depth (producer -> A loaders); the rows a conv needs
                                             // live in the TMEM row slots, the ring is TMA look-ahead only
constexpr int kSlotPix = 136;                // 130 px used (128 + 2 halo), rounded up to 8 px = 1024 B
constexpr int kSlotBytes = kSlotPix * 128;   // 17408, multiple of 1024
constexpr int kBoxPix = 130;
constexpr int kRowBytes = kBoxPix * 128;     // bytes one TMA row load delivers
constexpr int kAcc = 2;                      // TMEM accumulator buffers
constexpr int kARows = 4;                    // TMEM A-operand row slots (3 live + 1 being filled)
constexpr int kAColBase = 128;               // first TMEM column of the A region
constexpr int kAColsPerRow = 96;             // 3 dx copies x 32 columns
constexpr int kTmemCols = 512;
constexpr int kStageBytes = 128 * 128;       // one output row segment, bf16
constexpr int kThreads = 320;                // IN_TMA: producer, MMA, 4 epilogue, 4 loader warps
constexpr int kThreadsFused = 320 + 256;     // + two transform groups of 4 warps
constexpr int kThreadsTwoEpi = 320 + 128;    // EPI_SCALE_SKIP / EPI_RELU_STATS (IN_TMA): + a second epilogue group (warps 10-13)
constexpr const char* kDefaultL2Policy = "nnnnnn";  // see conv3x3_c64_tc(): overridden by DFIR_L2_POLICY
constexpr int kHlSlots = 3;                  // EPI_SCALE_SKIP_HL: stream-tile buffers per epilogue warp (1 in use, 2 in flight)
constexpr int kHlItemBytes = 4096;           // one tile: 16 px x 64 ch bf16 hi (2 KB) + lo (2 KB)
constexpr int kHl8Slots = 2;                 // EPI_SCALE_SKIP_HL8: ONE tile per warp and row (its 32 pixels: hi 4 KB + 8-bit lo 2 KB),
constexpr int kHl8ItemBytes = 6144;          //   two buffers: issuing a TMA operation costs the warp ~200 clk, so half as many
constexpr int kMaxBandImages = 8;            // a CTA's row band may touch at most this many images (IN_FUSED)
constexpr int kAttnScratchFloats = 64 + 64 + 512 + 1024 + 4;  // attention scratch of one epilogue group
constexpr int kCaStageFloats = 704;          // QCALayer parameter blobs up to this size are staged in shared memory

template <int NT>
struct SmemLayout {
  static constexpr int w_bytes = 9 * NT * 128;
  static constexpr int off_w = 0;
  static constexpr int off_ring = off_w + ((w_bytes + 1023) / 1024) * 1024;
  static constexpr int off_stage = off_ring + kSlots * kSlotBytes;
  // EPI_SCALE_SKIP: fp32 skip rows (cp.async targets): [half 0 | half 1] x [epilogue group] x [128 px][32 ch]
  static constexpr int off_skip = off_stage + 2 * kStageBytes;
  static constexpr int off_bias = off_skip + 2 * 128 * 64 * 4;
  static constexpr int off_pool = off_bias + 64 * 4;
  // Attention scratch (per epilogue group: y[64] s[64] attr[512] tmp[1024]) and the staged attention parameters are
  // only touched in the kernel prologue, before the first accumulator is drained: they alias the output staging tiles.
  static constexpr int off_attn = off_stage;
  static constexpr int off_cap = off_attn + 2 * kAttnScratchFloats * 4;
  static_assert(off_cap + kCaStageFloats * 4 <= off_stage + 2 * kStageBytes, "prologue scratch must fit the staging tiles");
  static constexpr int off_svec = off_pool + 12 * 64 * 4;  // (2 groups x [4 warp sums | first column | last column] x 64)
  static constexpr int off_bars = off_svec + kMaxBandImages * 64 * 4;
  static constexpr int n_bars = 2 * kSlots + kARows + 2 * kAcc + 1 + 8 * (kHlSlots > kHl8Slots ? kHlSlots : kHl8Slots);
  // EPI_SCALE_SKIP_HL: the 96 KB of the staging tiles + skip buffers hold, slot-major, kHlSlots x 8 warps x (2 KB hi +
  // 2 KB lo) stream tiles of 16 pixels; the prologue scratch aliases the last slot (first used after the prologue)
  static constexpr int off_hl = off_stage;
  static constexpr int hl_scratch_shift = (kHlSlots - 1) * 8 * kHlItemBytes;
  static constexpr int hl8_scratch_shift = (kHl8Slots - 1) * 8 * kHl8ItemBytes;
  static_assert(kHlSlots * 8 * kHlItemBytes <= 2 * kStageBytes + 2 * 128 * 64 * 4, "stream tiles must fit the epilogue buffers");
  static_assert(kHl8Slots * 8 * kHl8ItemBytes <= 2 * kStageBytes + 2 * 128 * 64 * 4, "stream tiles must fit the epilogue buffers");
  static_assert(2 * kAttnScratchFloats * 4 + kCaStageFloats * 4 <= 8 * kHl8ItemBytes, "prologue scratch must fit one slot");
  static constexpr int off_tmem = off_bars + n_bars * 8;
  static constexpr int total = off_tmem + 16;
};

// butterfly transpose-reduce over 32 values: on entry lane l holds v[0..31] (channel values of its pixel,
// already masked); on exit v[0] holds the sum over the warp's 32 pixels of channel l.
__device__ __forceinline__ void warp_channel_sums32(float (&v)[32], int lane) {
#pragma unroll
  for (int step = 0; step < 5; ++step) {
    const int n = 32 >> step;  // values held before this step
    const int mask = 16 >> step;
    const bool upper = (lane & mask) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = upper ? v[i] : v[i + n / 2];
      const float keep = upper ? v[i + n / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
    }
  }
}

// prmt.b32 with the full 4-bit selectors (bit 3 of a selector replicates the sign of the selected byte; __byte_perm only
// honours the low 3 bits)
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}

// shared-memory accesses by 32-bit shared-window address (the update loop of the stream epilogue: through generic pointers
// the compiler rebuilt the shared window base - S2UR SR_CgaCtaId + ULEA - in front of every group of accesses)
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

}  // namespace

// Epilogues that one group of four warps cannot drain at the MMA pace get a second group (warps 10-13) working on the
// alternate accumulator: the fp32 stream epilogue, and the ones with per-row channel sums (62 shuffles per thread and
// row) or a global-memory mask read per pixel.
template <int EPI, int INMODE>
constexpr bool two_epilogue_groups() {
  return INMODE == IN_TMA && EPI != EPI_TAIL_NCHW;
}

template <int EPI, int INMODE>
constexpr int conv_threads() {
  return INMODE == IN_FUSED ? kThreadsFused : (two_epilogue_groups<EPI, INMODE>() ? kThreadsTwoEpi : kThreads);
}

template <int NT, int EPI, int INMODE>
__global__ void __launch_bounds__(conv_threads<EPI, INMODE>(), 1)
conv3x3_c64_tc_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_out,
                      const __grid_constant__ ConvHlMaps hl, ConvTcArgs a) {
  using L = SmemLayout<NT>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms (TMA and UMMA) need 1024-byte alignment of every tile base
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* wsm = smem + L::off_w;
  uint8_t* ring = smem + L::off_ring;
  uint8_t* stage = smem + L::off_stage;
  float* skipbuf = reinterpret_cast<float*>(smem + L::off_skip);
  float* bias_s = reinterpret_cast<float*>(smem + L::off_bias);
  float* pool_s = reinterpret_cast<float*>(smem + L::off_pool);
  constexpr bool kHL = EPI == EPI_SCALE_SKIP_HL || EPI == EPI_SCALE_SKIP_HL8;  // hi + lo stream epilogue
  constexpr bool kLo8 = EPI == EPI_SCALE_SKIP_HL8;                             // ... with the 8-bit lo plane
  constexpr int kTilePx = kLo8 ? 32 : 16;                                      // pixels of a stream tile
  constexpr int kHlHiBytes = kTilePx * 128;                                    // hi part of a stream tile
  constexpr int kHlLoBytes = kLo8 ? kTilePx * 64 : kTilePx * 128;              // lo part of a stream tile
  constexpr bool kScaleSkip = EPI == EPI_SCALE_SKIP || kHL;
  // Descending traversal (hi + lo stream, nseg == 1): the pipeline runs on VIRTUAL coordinates (image B-1-b, row H-1-y,
  // kernel row 2-dy: a vertically flipped problem, ascending); only the addresses at the edges are mirrored.
  const bool flip = kHL && a.flip != 0;
  // Register reallocation (setmaxnreg): the kernel launches with 128 registers per thread (448 threads); the four loader
  // warps (physical warps 4-7, one aligned warpgroup) give registers back to the CTA's pool, the two epilogue groups
  // (physical warps 0-3 and 8-11) take them.  The pool is the CTA's own allocation (registers the SM has left over are NOT
  // in it: a first version that counted on them deadlocked in setmaxnreg.inc), so the books must balance inside the CTA:
  // 4 x 32 x (128 - 64) released = 8192 = 8 x 32 x (160 - 128) taken.
  constexpr bool kRegRealloc = two_epilogue_groups<EPI, INMODE>();
  constexpr int kLoaderRegs = 64, kEpilogueRegs = 160;
  constexpr int kHS = kLo8 ? kHl8Slots : kHlSlots;            // stream-tile buffers per epilogue warp
  constexpr int kHI = kLo8 ? kHl8ItemBytes : kHlItemBytes;    // bytes of one buffer
  constexpr int kScratchShift = kHL ? (kLo8 ? L::hl8_scratch_shift : L::hl_scratch_shift) : 0;
  float* attn_s = reinterpret_cast<float*>(smem + L::off_attn + kScratchShift);
  float* svec_s = reinterpret_cast<float*>(smem + L::off_svec);
  float* cap_s = reinterpret_cast<float*>(smem + L::off_cap + kScratchShift);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::off_bars);
  uint64_t* full = bars;                                   // ring slot filled      (producer/transform -> loaders)
  uint64_t* empty = full + kSlots;                         // ring slot drained     (loaders -> producer/transform)
  uint64_t* aempty = empty + kSlots;                       // TMEM A row slot free  (MMA commit -> loaders)
  uint64_t* tfull = aempty + kARows;                       // accumulator ready     (MMA commit -> epilogue)
  uint64_t* go = tfull + kAcc;                             // output row may start: its bottom A row is in TMEM (4 loader
                                                           // warps) and its accumulator is drained (4 epilogue warps)
  uint64_t* wbar = go + kAcc;                              // weights landed
  uint64_t* sbar = wbar + 1;                               // EPI_SCALE_SKIP_HL: stream tile landed, [8 warps][kHlSlots]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + L::off_tmem);

  // Physical warp order: [epilogue group 0: 0-3][A loaders: 4-7][epilogue group 1 or input transform: 8..][TMA producer]
  // [MMA issuer].  The warp scheduler of an SM sub-partition (warp id % 4) favours the highest warp id (measured on
  // sm_10x, B300_MICROARCH.md): the MMA thread must never lose an issue slot to an epilogue or loader warp of its
  // sub-partition, so it sits in the last warp and the producer in the one before.  `warp` / `tid` below are the LOGICAL
  // role indices the rest of the kernel is written in (0 producer, 1 MMA, 2-5 epilogue, 6-9 loaders, 10.. second
  // epilogue group / transform); tensor-memory lane quarters follow the physical warp id (`pwarp & 3`).
  const int pwarp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int warp = pwarp == nwarps - 2 ? 0 : (pwarp == nwarps - 1 ? 1 : pwarp + 2);
  const int lane = threadIdx.x & 31;
  const int tid = warp * 32 + lane;

  // work partition: G = ncols * H row segments, split evenly over the grid
  const int H = a.H;
  const int nseg = a.nseg;
  const long long G = static_cast<long long>(a.B) * nseg * H;
  const int g0 = static_cast<int>(G * blockIdx.x / gridDim.x);
  const int g1 = static_cast<int>(G * (blockIdx.x + 1) / gridDim.x);
  const int Hp = H + 2;
  // padded row index of output row g: col*(H+2) + y + 1
  auto padded = [&](int g) { return (g / H) * Hp + (g % H) + 1; };
  const int pr_first = (g0 < g1) ? padded(g0) - 1 : 0;
  const int n_last = (g0 < g1) ? padded(g1 - 1) + 1 - pr_first : -1;  // last ring sequence index

  if (threadIdx.x == 0) {
    if (INMODE == IN_TMA) prefetch_tmap(&tmap_in);
    if (EPI != EPI_TAIL_NCHW && !kScaleSkip) prefetch_tmap(&tmap_out);
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(&full[i], INMODE == IN_FUSED ? 128 : 1);
      mbar_init(&empty[i], 4);
    }
    for (int i = 0; i < kARows; ++i) mbar_init(&aempty[i], 1);
    for (int i = 0; i < kAcc; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&go[i], 8);
    }
    mbar_init(wbar, 1);
    if constexpr (kHL) {
      for (int i = 0; i < 8 * kHS; ++i) mbar_init(&sbar[i], 1);
      for (int i = 0; i < 4; ++i) prefetch_tmap(&hl.m[i]);
    }
    fence_barrier_init();
  }
  if (tid >= 64 && tid < 64 + NT) bias_s[tid - 64] = a.bias != nullptr ? a.bias[tid - 64] : 0.f;
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_holder);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  // Programmatic dependent launch: the next conv of the chain may become resident as soon as this CTA retires; its
  // barrier set-up, TMEM allocation and weight load then overlap the tail of this grid.  Everything that reads what
  // the previous kernel wrote (activation rows, statistics, skip rows) sits behind grid_dep_wait().
  if (threadIdx.x == 0) grid_dep_launch();
  // The activations of one RCAB (~0.5 GB at 32 images) flush the 126 MB L2 between two uses of a layer's weights, so
  // each conv would start with 148 CTAs missing on the same 72 KB.  The first CTAs pull the next conv's weights into
  // L2 while this one runs.
  if (a.next_w != nullptr && warp == 2) {
    const int nlines = a.next_w_bytes >> 7;
    for (int line = blockIdx.x * 32 + lane; line < nlines; line += gridDim.x * 32)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const uint8_t*>(a.next_w) + static_cast<size_t>(line) * 128));
  }
  // Timing experiments (wrong results) and progress probes are compiled in only with -DDFIR_PROBES (make PROBES=1):
  // in the shipped kernel they are constant-false, so neither their branches nor their code reach the hot loops.
#ifdef DFIR_PROBES
  const bool probe = (a.debug_probe & 1) != 0 && blockIdx.x == 0 && lane == 0;
  const bool exp_skip_store = (a.debug_probe & 2) != 0;  // timing experiments only (wrong results)
  const bool exp_one_copy = (a.debug_probe & 4) != 0;
  const bool exp_no_copy = (a.debug_probe & 8) != 0;
  const bool exp_no_epi = (a.debug_probe & 32) != 0;
  const bool exp_n192 = (a.debug_probe & 64) != 0;   // 12 MMAs of N=192 per row (timing only)
  const bool exp_n128 = (a.debug_probe & 128) != 0;  // 18 MMAs of N=128 per row (timing only)
  const bool exp_no_skipld = (a.debug_probe & 256) != 0;
  const bool exp_no_f32st = (a.debug_probe & 512) != 0;
  const bool exp_no_bfst = (a.debug_probe & 1024) != 0;
  const bool exp_no_pf = (a.debug_probe & 2048) != 0;      // no L2 prefetch of the skip rows
  const bool exp_no_tile = (a.debug_probe & 4096) != 0;    // scale+skip epilogue: TMEM reads and barriers only
  const bool exp_no_pass = (a.debug_probe & 8192) != 0;    // scale+skip epilogue: no coalesced pass
  const bool trace_on = (a.debug_probe & 32768) != 0 && blockIdx.x == gridDim.x / 2;
  // (PROBES build) bit 262144: the columns of epilogue group 1 record the tile loop of group 0 instead (HL epilogues)
  const bool tile_probe = (a.debug_probe & 262144) != 0;
  const bool exp_no_lo = (a.debug_probe & 524288) != 0;   // HL8 update without the lo-plane shared-memory accesses (wrong results)
  const bool exp_no_hi = (a.debug_probe & 1048576) != 0;  // ... without the hi-plane accesses
#else
  constexpr bool probe = false, exp_skip_store = false, exp_one_copy = false, exp_no_copy = false, exp_no_epi = false,
                 exp_no_skipld = false, exp_no_f32st = false, exp_no_bfst = false,
                 exp_no_pf = false, exp_no_tile = false, exp_no_pass = false, tile_probe = false, exp_no_lo = false, exp_no_hi = false;
#endif

  if (g0 < g1) {
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (elect_one()) {
        mbar_arrive_expect_tx(wbar, L::w_bytes);
        bulk_load_1d(wsm, a.wpacked, L::w_bytes, wbar);  // weights are not produced by the previous kernel
        grid_dep_wait();
        if constexpr (INMODE == IN_TMA) {
          int col = pr_first / Hp, yy = pr_first % Hp - 1;  // advanced without divisions
          int cseg = col % nseg, cimg = col / nseg;
          for (int n = 0; n <= n_last; ++n) {
            const int slot = n % kSlots;
            if (probe) g_dfir_progress[0] = n + 1;
            mbar_wait(&empty[slot], ((n / kSlots) & 1) ^ 1, 1);
            DFIR_TRACE(0, n);  // producer: ring slot free, TMA load of row n issued
            mbar_arrive_expect_tx(&full[slot], kRowBytes);
            const int yy_a = flip ? H - 1 - yy : yy, cimg_a = flip ? a.B - 1 - cimg : cimg;
            if (a.use_hints)
              tma_load_4d_hint(ring + slot * kSlotBytes, &tmap_in, &full[slot], a.cin_off, cseg * 128 - 1, yy_a, cimg_a, a.pol_in);
            else
              tma_load_4d(ring + slot * kSlotBytes, &tmap_in, &full[slot], a.cin_off, cseg * 128 - 1, yy_a, cimg_a);
            if (++yy > H) {  // past the bottom pad row: next column segment / image
              yy = -1;
              if (++cseg == nseg) {
                cseg = 0;
                ++cimg;
              }
            }
          }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      // Measured (tools/mma_probe.cu, tools/trace_conv.py; profiles/r02_mma_issue.md): a cta_group::1 TS-mode 128x64x16
      // MMA stream runs at its 32-clk floor, but the MMA queue is only ~2 instructions deep, so every clock the issuing
      // thread spends on anything else (an mbarrier test costs 90-230 clk even when the phase completed long ago, a
      // divergent-branch reconvergence, integer divisions, bursts of R2UR) is tensor-pipe idle time.  Hence: ONE elected
      // thread runs the whole loop, one mbarrier wait per output row (`go[acc]`: the row's bottom A row has been copied
      // into tensor memory by the four loader warps AND the accumulator has been drained by the four epilogue warps that
      // own it: 8 arrivals), the 36 MMAs and the two commits of a row in one basic block, no division.
      if (elect_one()) {
        constexpr uint32_t idesc = make_idesc_bf16_f32(128, NT);
        const uint64_t db_base = make_sw128_kmajor_desc(smem_u32(wsm), 1024, 0);
        uint64_t tap_row[3];  // weight descriptors of the kernel row that meets the window's top / centre / bottom A row
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
          tap_row[dy] = db_base + static_cast<uint32_t>((((flip ? 2 - dy : dy) * 3) * NT * 128) >> 4);
        mbar_wait(wbar, 0, 2);
        // Row state, computed one row ahead: the queue of issued-but-unfinished MMAs is only ~4 deep (~130 clk of tensor
        // work), so the ~30 integer instructions and the barrier test (~100 clk) that separate two rows are issued in
        // front of the LAST eight MMAs of the previous row, while those MMAs wait for queue slots anyway.
        int y_cur = g0 % H;
        int nc = padded(g0) - pr_first;  // sequence index of the centre row
        auto row_cols = [&](int ncv, uint32_t (&cols)[3]) {
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) cols[dy] = tmem_base + kAColBase + ((ncv - 1 + dy) & (kARows - 1)) * kAColsPerRow;
        };
        uint32_t a_col[3];
        row_cols(nc, a_col);
        mbar_wait(&go[0], 0, 3);
        for (int g = g0, it = 0; g < g1; ++g, ++it) {
          const bool img_end = y_cur == H - 1;  // the next output row starts a new image (or column segment)
          const bool last = g + 1 == g1;
          const int acc = it & 1;
          const uint32_t d_tmem = tmem_base + acc * NT;
          uint64_t* const rel_bar = &aempty[(nc - 1) & (kARows - 1)];
          if (probe) g_dfir_progress[1] = it + 1;
          tcgen05_fence_after();
          DFIR_TRACE1(5, it);  // MMA thread: row starts
          auto issue = [&](int first, int count) {  // MMAs [first, first + count) of the row, in (dy, dx, k) order
#pragma unroll
            for (int j = first; j < first + count; ++j) {
              const int dy = j / 12, dx = (j / 4) % 3, k = j & 3;
              const uint64_t db = tap_row[dy] + static_cast<uint32_t>((dx * NT * 128 + k * 32) >> 4);
              umma_f16_ts(d_tmem, a_col[dy] + dx * 32 + k * 8, db, idesc, j != 0 ? 1u : 0u);
            }
          };
#ifdef DFIR_PROBES
          if (exp_n192 || exp_n128) {
            if constexpr (NT == 64) {
              constexpr uint32_t id192 = make_idesc_bf16_f32(128, 192);
              constexpr uint32_t id128 = make_idesc_bf16_f32(128, 128);
              const int reps = exp_n192 ? 1 : 2;
              for (int rpt = 0; rpt < reps; ++rpt)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const uint64_t db = db_base + static_cast<uint32_t>(((dx * 3) * NT * 128 + k * 32) >> 4);
                    umma_f16_ts(tmem_base, a_col[1] + dx * 32 + k * 8, db, exp_n192 ? id192 : id128, 1u);
                  }
              if (!exp_n192)
                for (int k = 0; k < 6; ++k)
                  umma_f16_ts(tmem_base, a_col[1] + k * 8, db_base + static_cast<uint32_t>((k * 32) >> 4), id128, 1u);
            }
            umma_commit(rel_bar);
          } else
#endif
          {
            issue(0, 12);
            // the top row is dead as soon as the dy = 0 taps have executed: hand its TMEM slot back now so that the
            // loaders refill it while the remaining 24 MMAs of this row run
            umma_commit(rel_bar);
            issue(12, 16);
          }
          // ---- next row's state and barrier, behind the MMAs queued so far
          const int nc_next = nc + (img_end ? 3 : 1);
          uint32_t a_col_next[3];
          row_cols(nc_next, a_col_next);
          bool next_ready = true;
          if (!last) next_ready = mbar_try_wait(&go[acc ^ 1], ((it + 1) >> 1) & 1);
#ifdef DFIR_PROBES
          if (!(exp_n192 || exp_n128))
#endif
            issue(28, 8);
          umma_commit(&tfull[acc]);
          DFIR_TRACE1(7, it);  // MMA thread: 36 MMAs + commits issued
          if (img_end || last) {
            // at an image boundary the next centre also skips the last row and the bottom pad row of the finished
            // image — keeping them would deadlock the 4-slot TMEM row ring
            umma_commit(&aempty[nc & (kARows - 1)]);
            umma_commit(&aempty[(nc + 1) & (kARows - 1)]);
          }
          if (!next_ready) {
            DFIR_TRACE1(4, it);  // MMA thread: the early test failed, blocking wait for the next row
            mbar_wait(&go[acc ^ 1], ((it + 1) >> 1) & 1, 3);
          }
          nc = nc_next;
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) a_col[dy] = a_col_next[dy];
          y_cur = img_end ? 0 : y_cur + 1;
        }
      }
      __syncwarp();
    } else if (warp >= 6 && warp < 10) {
      // ===================== A loaders: smem ring row -> 3 dx-shifted copies in TMEM =====================
      if constexpr (kRegRealloc) setmaxnreg_dec<kLoaderRegs>();
      const int q = pwarp & 3;       // TMEM lane quarter this warp may access
      const int m = q * 32 + lane;   // output pixel = TMEM lane
      // Row n is the LAST A row some output row waits for (its bottom row) iff it is at least the third row of its
      // padded image and of this band; the rows complete in order, so that output row's `go` barrier gets this warp's
      // arrival here and nothing else has to be signalled to the MMA thread.
      int p_img = pr_first % Hp;     // padded row index of ring row n inside its image: 0 = top pad row .. H + 1 = bottom pad
      int n_out = 0;                 // output rows of this band signalled so far
      for (int n = 0; n <= n_last; ++n) {
        const int slot = n % kSlots;
        const int as = n % kARows;
        if (probe) g_dfir_progress[4 + q] = n + 1;
        mbar_wait(&full[slot], (n / kSlots) & 1, 5);
        if (q == 0) DFIR_TRACE(1, n);  // loader: row n has landed in the ring
        mbar_wait(&aempty[as], ((n / kARows) & 1) ^ 1, 6);
        if (q == 0) DFIR_TRACE(2, n);  // loader: TMEM row slot free
        tcgen05_fence_after();
        const uint8_t* srow = ring + slot * kSlotBytes;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + kAColBase + as * kAColsPerRow;
        if (exp_no_copy) {
        } else if (exp_one_copy) {
          const int p = m + 1;
          const uint4* src = reinterpret_cast<const uint4*>(srow + p * 128);
          uint32_t v[32];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 t = src[c ^ (p & 7)];
            v[4 * c + 0] = t.x; v[4 * c + 1] = t.y; v[4 * c + 2] = t.z; v[4 * c + 3] = t.w;
          }
          tmem_st_32x32b_x32(t_row, v);
          tmem_st_32x32b_x32(t_row + 32, v);
          tmem_st_32x32b_x32(t_row + 64, v);
        } else {
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int p = m + dx;  // ring pixel feeding output pixel m for this tap column
          const uint4* src = reinterpret_cast<const uint4*>(srow + p * 128);
          uint32_t v[32];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 t = src[c ^ (p & 7)];
            v[4 * c + 0] = t.x;
            v[4 * c + 1] = t.y;
            v[4 * c + 2] = t.z;
            v[4 * c + 3] = t.w;
          }
          tmem_st_32x32b_x32(t_row + dx * 32, v);
        }
        }
        tmem_st_wait();
        if (q == 0) DFIR_TRACE(3, n);  // loader: the three TMEM copies of row n are complete
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (n >= 2 && p_img >= 2) mbar_arrive(&go[n_out & 1]);
          mbar_arrive(&empty[slot]);
        }
        if (n >= 2 && p_img >= 2) ++n_out;
        if (++p_img == Hp) p_img = 0;
      }
    } else if (warp >= 10 && !two_epilogue_groups<EPI, INMODE>()) {
      // ===================== fused input transform (IN_FUSED only; warps 10..17) =====================
      if constexpr (INMODE == IN_FUSED) {
        grid_dep_wait();
        const int tt = tid - kThreads;  // 0..255
        // ---- prologue (overlaps the weight load): attention vectors of the images this band touches.
        //      s[b] = CA_style(mean over pixels of r_b, attributes[b]) * meta_scale[b]; the pooled mean is
        //      rebuilt from the per-row sums in a fixed order, so it is bit-identical in every CTA.
        const int rows_img = nseg * H;
        const int b_first = g0 / rows_img, b_last = (g1 - 1) / rows_img;
        {
          float* y_s = attn_s;
          float* s_s = attn_s + 64;
          float* attr_s = attn_s + 128;
          float* tmp = attn_s + 128 + 512;
          const NamedGroup grp{tt, 256, 5};
          for (int b = b_first; b <= b_last; ++b) {
            float* s_img = svec_s + (b - b_first) * 64;
            if (a.ca_style == DFIR_STYLE_NONE) {
              if (tt < 64) s_img[tt] = a.res_scale * (a.sq != nullptr ? a.sq[static_cast<size_t>(b) * 64 + tt] : 1.f);
              grp.sync();
              continue;
            }
            const int cq = tt & 15, rg = tt >> 4;  // channel quad, row group (16 groups)
            const float4* pr = reinterpret_cast<const float4*>(a.pool_rows + static_cast<size_t>(b) * rows_img * 64) + cq;
            float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int row = rg; row < rows_img; row += 16) {
              const float4 v = pr[static_cast<size_t>(row) * 16];
              acc4.x += v.x; acc4.y += v.y; acc4.z += v.z; acc4.w += v.w;
            }
            reinterpret_cast<float4*>(tmp)[rg * 16 + cq] = acc4;
            for (int i = tt; i < a.ca_A; i += 256) attr_s[i] = a.attributes[static_cast<size_t>(b) * a.ca_A + i];
            grp.sync();
            if (tt < 64) {
              float t = 0.f;
#pragma unroll
              for (int k = 0; k < 16; ++k) t += tmp[k * 64 + tt];
              y_s[tt] = t / (static_cast<float>(H) * static_cast<float>(a.W));
            }
            grp.sync();
            attn_vector(grp, a.ca_style, a.ca_params, 64, a.ca_R, a.ca_M, attr_s, y_s, s_s, tmp);
            if (tt < 64) s_img[tt] = s_s[tt] * (a.sq != nullptr ? a.sq[static_cast<size_t>(b) * 64 + tt] : 1.f);
            grp.sync();
          }
        }
        // ---- row transform: two groups of 128 threads alternate ring rows.  Thread = (4-channel group c4,
        //      pixel p0 + 8j): a warp covers 2 pixels x 64 channels per instruction, so the fp32 stream is
        //      read and written in fully coalesced 512 B requests and r in 256 B requests.
        const int tg = tt >> 7;
        const int tl = tt & 127;
        const int c4 = tl & 15, p0 = tl >> 4;
        const int pc_first = padded(g0), pc_last = padded(g1 - 1);
        const uint32_t sw_chunk = static_cast<uint32_t>(c4 >> 1), sw_half = static_cast<uint32_t>(c4 & 1) * 8u;
        int cur_b = -1;
        float4 sc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int n = tg; n <= n_last; n += 2) {
          const int pr = pr_first + n;
          const int slot = n % kSlots;
          const int col = pr / Hp;
          const int yy = pr % Hp - 1;
          const int b = col / nseg;
          const int seg = col % nseg;
          const bool row_ok = yy >= 0 && yy < H;
          const bool owned = row_ok && pr >= pc_first && pr <= pc_last && a.xout_f32 != nullptr;
          if (row_ok && b != cur_b) {
            sc = *reinterpret_cast<const float4*>(svec_s + (b - b_first) * 64 + c4 * 4);
            cur_b = b;
          }
          uint8_t* srow = ring + slot * kSlotBytes;
          const size_t row_e = (static_cast<size_t>(b) * H + (row_ok ? yy : 0)) * a.W * 64;
          const int xbase = seg * 128 - 1;
          bool waited = false;
#pragma unroll 1
          for (int j0 = 0; j0 < 17; j0 += 9) {  // pixels p0 + 8j, j = 0..16, in two batches of loads
            uint2 rr[9];
            float4 xx[9];
            bool ok[9];
#pragma unroll
            for (int u = 0; u < 9; ++u) {
              const int p = p0 + 8 * (j0 + u);
              const int x = xbase + p;
              ok[u] = row_ok && (j0 + u) < 17 && p < kBoxPix && x >= 0 && x < a.W;
              if (ok[u]) {
                const size_t e = row_e + static_cast<size_t>(x) * 64 + c4 * 4;
                rr[u] = *reinterpret_cast<const uint2*>(a.r_bf16 + e);
                xx[u] = *reinterpret_cast<const float4*>(a.xin_f32 + e);
              }
            }
            if (!waited) {  // the ring slot is only needed once the loads are in flight
              mbar_wait(&empty[slot], ((n / kSlots) & 1) ^ 1, 7);
              waited = true;
            }
#pragma unroll
            for (int u = 0; u < 9; ++u) {
              const int p = p0 + 8 * (j0 + u);
              if ((j0 + u) >= 17 || p >= kBoxPix) continue;
              uint2 pk = make_uint2(0u, 0u);  // zero padding (rows -1 / H, columns -1 / W)
              if (ok[u]) {
                const float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rr[u].x));
                const float2 f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rr[u].y));
                float4 o;
                o.x = fmaf(f0.x, sc.x, xx[u].x);
                o.y = fmaf(f0.y, sc.y, xx[u].y);
                o.z = fmaf(f1.x, sc.z, xx[u].z);
                o.w = fmaf(f1.y, sc.w, xx[u].w);
                if (owned && p >= 1 && p <= 128)
                  *reinterpret_cast<float4*>(a.xout_f32 + row_e + static_cast<size_t>(xbase + p) * 64 + c4 * 4) = o;
                pk.x = pack_bf16x2(o.x, o.y);
                pk.y = pack_bf16x2(o.z, o.w);
              }
              *reinterpret_cast<uint2*>(srow + p * 128 + ((sw_chunk ^ static_cast<uint32_t>(p & 7)) << 4) + sw_half) = pk;
            }
          }
          mbar_arrive(&full[slot]);
        }
      }
    } else {
      // ===================== epilogue (warps 2..5; EPI_SCALE_SKIP / EPI_RELU_STATS with TMA input: a second group,
      // warps 10..13) ======
      // EPI_SCALE_SKIP moves 10 B per element through the epilogue and is latency bound in one group (measured
      // 3 us per row against 1.1 us of MMAs), so two groups take alternate rows (= alternate accumulators).
      // EPI_RELU_STATS: the per-row channel sums (62 shuffles per thread and row) take one group 1.76 us per row
      // against 1.3 us of the MMA pipeline (conv1 48.7 us vs 38 us without the statistics), same remedy.
      constexpr bool kTwoEpi = two_epilogue_groups<EPI, INMODE>();
      constexpr int kEpiGroups = kTwoEpi ? 2 : 1;
      if constexpr (kRegRealloc) setmaxnreg_inc<kEpilogueRegs>();
      const int egrp = (kTwoEpi && warp >= 10) ? 1 : 0;
      const int q = pwarp & 3;         // TMEM lane quarter this warp may read
      const int m = q * 32 + lane;     // pixel within the 128-px row segment
      const int et = egrp ? tid - 320 : tid - 64;  // 0..127 within the group
      // both accumulators start out drained: the first use of each must not wait for an epilogue arrival on `go`
      if (lane == 0) {
        if (kEpiGroups == 1) {
          mbar_arrive(&go[0]);
          mbar_arrive(&go[1]);
        } else {
          mbar_arrive(&go[egrp]);
        }
      }
      const int rows_img_e = nseg * H;
      const int bimg_first = g0 / rows_img_e;
      int cur_img = -1;
      long long fx_t = 0, fx_c0 = 0, fx_cl = 0;  // EPI_RELU_STATS fixed-point mode: this thread's channel, current image
      // EPI_RELU_STATS_W: the sums of this thread's 16 channels over its pixels of the current image, in 2^-12 units
      int w_t[16], w_c[16], w_cw = 0;   // w_c: this lane's pixel of the first (w_cw = 1) or last (2) column, if it owns one
      if constexpr (EPI == EPI_RELU_STATS_W) {
#pragma unroll
        for (int j = 0; j < 16; ++j) w_t[j] = w_c[j] = 0;
      }
      // ---- EPI_SCALE_SKIP_HL: every epilogue warp streams the residual tiles of its own 32 pixels through kHlSlots
      // private 4 KB buffers (TMA load -> in-place update -> TMA store), two tiles of 16 pixels per output row, loads
      // issued two tiles ahead.  No barrier other than the tile's own mbarrier: the warps never wait for each other.
      const int hl_w = egrp * 4 + q;                       // buffer column of this warp
      uint8_t* const hl_base = smem + L::off_hl + hl_w * kHI;
      const uint32_t hl_sbase = smem_u32(hl_base);
      uint64_t* const hl_bar = sbar + hl_w * kHS;
      int hl_ly = (g0 + egrp) % H, hl_lcol = (g0 + egrp) / H;  // load cursor: next row whose tiles are requested
      int hl_lb = hl_lcol / nseg, hl_lseg = hl_lcol % nseg;
      int hl_lg = g0 + egrp, hl_lhalf = 0, hl_lslot = 0;
      auto hl_issue = [&]() {  // lane 0: request the next tile (if any) into the next slot
        if (hl_lg < g1) {
          uint8_t* dst = hl_base + hl_lslot * (8 * kHI);
          const int xs = hl_lseg * 128 + q * 32 + hl_lhalf * kTilePx;
          mbar_arrive_expect_tx(&hl_bar[hl_lslot], kHlHiBytes + kHlLoBytes);
          const int ya = flip ? H - 1 - hl_ly : hl_ly, ba = flip ? a.B - 1 - hl_lb : hl_lb;
          tma_load_4d(dst, &hl.m[0], &hl_bar[hl_lslot], 0, xs, ya, ba);
          if (a.use_hints) tma_load_4d_hint(dst + kHlHiBytes, &hl.m[1], &hl_bar[hl_lslot], 0, xs, ya, ba, a.pol_skip);
          else tma_load_4d(dst + kHlHiBytes, &hl.m[1], &hl_bar[hl_lslot], 0, xs, ya, ba);
          // A/B (DFIR_DEBUG_PROBE bit 65536): L2 prefetch of the tile two rows of this warp further down.  Measured: 62.2 ->
          // 67.9 us per launch — the kernel is HBM-bandwidth bound, not latency bound; default off.
          if ((a.debug_probe & 65536) != 0 && hl_ly + 2 * kEpiGroups < H && hl_lg + 2 * kEpiGroups < g1) {
            const int yp = flip ? H - 1 - (hl_ly + 2 * kEpiGroups) : hl_ly + 2 * kEpiGroups;
            tma_prefetch_4d(&hl.m[0], 0, xs, yp, ba);
            tma_prefetch_4d(&hl.m[1], 0, xs, yp, ba);
          }
        }
        if (++hl_lslot == kHS) hl_lslot = 0;
        if (++hl_lhalf == 32 / kTilePx) {
          hl_lhalf = 0;
          hl_lg += kEpiGroups;
          hl_ly += kEpiGroups;
          while (hl_ly >= H) {
            hl_ly -= H;
            if (++hl_lseg == nseg) {
              hl_lseg = 0;
              ++hl_lb;
            }
          }
        }
      };
      int hl_slot = 0, hl_phase = 0;                        // consume cursor
      float hl_s[16], hl_bs[16];                            // scale and bias * scale of this thread's 16 channels
      if constexpr (kScaleSkip) {
        // ---- pool-by-linearity: while the pipeline fills, the epilogue warps turn the sums of
        // t = relu(conv1(x)) left by the previous kernel into this block's attention vectors
        //   mean(conv2(t))[co] = b[co] + (1/HW) sum_{tap,ci} W[co][ci][tap] * S[tap][ci]
        // (W = this conv's weights, already on their way into shared memory), then QCALayer * meta scale.
        // The two epilogue groups take alternate images of the band.  Everything that does not depend on the previous
        // kernel (the block's attention parameters, staged into shared memory when they fit) is fetched before
        // griddepcontrol.wait.
        const float* cap = a.ca_params;
        if (a.epi_stats) {
          const int np = attn_param_count(a.ca_style, 64, a.ca_R, a.ca_M);
          if (np <= kCaStageFloats) {
            for (int i = et + egrp * 128; i < np; i += 128 * kEpiGroups) cap_s[i] = a.ca_params[i];
            cap = cap_s;
          }
        }
        grid_dep_wait();
        if (q == 0 && egrp == 0) DFIR_TRACE(8, 40);  // (PROBES build) prologue time stamps of epilogue group 0
        if constexpr (kHL) {
          if (lane == 0) {  // the first tiles travel while the attention vector is evaluated (8-bit lo: one tile = one row;
            hl_issue();     // the second buffer is the prologue's scratch until the barrier that ends the prologue)
            if (!kLo8) hl_issue();
          }
        }
        if (a.epi_stats) {
          const bool fixed_stats = a.epi_stats == 2;
          const int b_mine = bimg_first + egrp;                     // first image of this group
          const bool have_mine = b_mine <= (g1 - 1) / rows_img_e;
          float fxv[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          float sqv = 1.f;                                          // meta scale of the image (channel et)
          auto load_fixed = [&](int bimg) {  // threads 0..63: the nine fixed-point sums of one channel, the meta scale
            if (et < 64) {
              const long long* st = a.istats + static_cast<size_t>(bimg) * 576 + et;
              long long raw[9];
#pragma unroll
              for (int k = 0; k < 9; ++k) raw[k] = __ldcg(st + 64 * k);
              if (a.sq != nullptr) sqv = a.sq[static_cast<size_t>(bimg) * 64 + et];
#pragma unroll
              for (int k = 0; k < 9; ++k) fxv[k] = __ll2float_rn(raw[k]) * (1.f / 16777216.f);
            }
          };
          if (fixed_stats && have_mine) load_fixed(flip ? a.B - 1 - b_mine : b_mine);
          if constexpr (kTwoEpi) named_bar_sync(6, 256); else named_bar_sync(5, 128);  // cap_s complete
          if (q == 0 && egrp == 0) DFIR_TRACE(8, 41);
          float* scratch = attn_s + egrp * kAttnScratchFloats;
          float* y_s = scratch;
          float* s_s = scratch + 64;
          float* attr_s = scratch + 128;
          float* tmp = scratch + 128 + 512;  // 1024 floats
          const NamedGroup grp{et, 128, egrp ? 7 : 5};
          const int bimg_last = (g1 - 1) / rows_img_e;
          const int cq = et & 15, rg = et >> 4;  // channel quad, row group (8 groups: 2 per warp)
          const float inv_hw = 1.f / (static_cast<float>(H) * static_cast<float>(a.W));
          auto add4 = [](float4& d, const float4 v) { d.x += v.x; d.y += v.y; d.z += v.z; d.w += v.w; };
          auto fold = [&](float4 v) {  // + the other row group of this warp (lane ^ 16)
            v.x += __shfl_xor_sync(0xffffffffu, v.x, 16); v.y += __shfl_xor_sync(0xffffffffu, v.y, 16);
            v.z += __shfl_xor_sync(0xffffffffu, v.z, 16); v.w += __shfl_xor_sync(0xffffffffu, v.w, 16);
            return v;
          };
          mbar_wait(wbar, 0, 9);  // conv weights have landed in smem (generic-proxy reads below)
          if (q == 0 && egrp == 0) DFIR_TRACE(8, 42);
          for (int b = bimg_first + egrp; b <= bimg_last; b += kEpiGroups) {
            const int ba = flip ? a.B - 1 - b : b;  // the image the statistics, attributes and meta scale belong to
            // statistics source: per-row arrays written by conv1 (summed here in a fixed order), or the nine 64-bit fixed-point
            // sums per channel that conv1 accumulated with atomics (epi_stats == 2: one 72-byte read per thread, issued
            // ahead of everything else for the group's first image)
            float r0v = 0.f, rlv = 0.f, k00 = 0.f, k0w = 0.f, kh0 = 0.f, khw = 0.f;
            if (fixed_stats) {
              if (b != b_mine) load_fixed(ba);
            } else {
              if (et < 64 && a.sq != nullptr) sqv = a.sq[static_cast<size_t>(ba) * 64 + et];
            const float4* pr = reinterpret_cast<const float4*>(a.pool_rows + static_cast<size_t>(ba) * rows_img_e * 64) + cq;
            const float4* cf = reinterpret_cast<const float4*>(a.col_first + static_cast<size_t>(ba) * H * 64) + cq;
            const float4* cl = reinterpret_cast<const float4*>(a.col_last + static_cast<size_t>(ba) * H * 64) + cq;
            // edge rows / corners of the stats (threads 0..63, one channel each): issued ahead of the row loops so
            // that their latency overlaps, consumed after the first barrier
            if (et < 64) {
              const float* prs = a.pool_rows + static_cast<size_t>(ba) * rows_img_e * 64 + et;
              const float* cfs = a.col_first + static_cast<size_t>(ba) * H * 64 + et;
              const float* cls = a.col_last + static_cast<size_t>(ba) * H * 64 + et;
              r0v = prs[0];
              rlv = prs[static_cast<size_t>(H - 1) * 64];
              k00 = cfs[0]; k0w = cls[0];
              kh0 = cfs[static_cast<size_t>(H - 1) * 64]; khw = cls[static_cast<size_t>(H - 1) * 64];
            }
            float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f), c04 = t4, c14 = t4;
            {  // explicit batches of 8 independent 16-byte loads (fixed summation order)
              int row = rg;
              for (; row + 56 < rows_img_e; row += 64) {
                float4 v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = pr[static_cast<size_t>(row + 8 * k) * 16];
#pragma unroll
                for (int k = 0; k < 8; ++k) add4(t4, v[k]);
              }
              for (; row < rows_img_e; row += 8) add4(t4, pr[static_cast<size_t>(row) * 16]);
              int yy = rg;
              for (; yy + 24 < H; yy += 32) {
                float4 v[8];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  v[2 * k] = cf[static_cast<size_t>(yy + 8 * k) * 16];
                  v[2 * k + 1] = cl[static_cast<size_t>(yy + 8 * k) * 16];
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  add4(c04, v[2 * k]);
                  add4(c14, v[2 * k + 1]);
                }
              }
              for (; yy < H; yy += 8) {
                add4(c04, cf[static_cast<size_t>(yy) * 16]);
                add4(c14, cl[static_cast<size_t>(yy) * 16]);
              }
            }
            t4 = fold(t4); c04 = fold(c04); c14 = fold(c14);
            if ((lane & 16) == 0) {  // tmp: [3 sums][4 warps][64]
              const int w4 = et >> 5;
              reinterpret_cast<float4*>(tmp + (0 * 4 + w4) * 64)[cq] = t4;
              reinterpret_cast<float4*>(tmp + (1 * 4 + w4) * 64)[cq] = c04;
              reinterpret_cast<float4*>(tmp + (2 * 4 + w4) * 64)[cq] = c14;
            }
            }
            for (int i = et; i < a.ca_A; i += 128) attr_s[i] = a.attributes[static_cast<size_t>(ba) * a.ca_A + i];
            if (!fixed_stats) grp.sync();
            if (q == 0 && egrp == 0 && !fixed_stats) DFIR_TRACE(8, 43);
            float* S = tmp;  // [9][64], written once the partial sums have been consumed (fixed-point sums: no partial sums)
            float Sv[9];
            if (et < 64) {
              const int c = et;
              float T, C0, CL, R0, RL;
              if (fixed_stats) {
                T = fxv[0]; C0 = fxv[1]; CL = fxv[2]; R0 = fxv[3]; RL = fxv[4];
                k00 = fxv[5]; k0w = fxv[6]; kh0 = fxv[7]; khw = fxv[8];
              } else {
                T = (tmp[c] + tmp[64 + c]) + (tmp[128 + c] + tmp[192 + c]);
                C0 = (tmp[256 + c] + tmp[320 + c]) + (tmp[384 + c] + tmp[448 + c]);
                CL = (tmp[512 + c] + tmp[576 + c]) + (tmp[640 + c] + tmp[704 + c]);
                R0 = r0v; RL = rlv;
                for (int sg = 1; sg < nseg; ++sg) {  // images wider than one 128-px segment
                  const float* prs = a.pool_rows + static_cast<size_t>(ba) * rows_img_e * 64 + c;
                  R0 += prs[(static_cast<size_t>(sg) * H) * 64];
                  RL += prs[(static_cast<size_t>(sg) * H + (H - 1)) * 64];
                }
              }
#pragma unroll
              for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                  const float rowx = dy == 0 ? RL : (dy == 2 ? R0 : 0.f);
                  const float colx = dx == 0 ? CL : (dx == 2 ? C0 : 0.f);
                  float corner = 0.f;
                  if (dy == 0 && dx == 0) corner = khw;
                  if (dy == 0 && dx == 2) corner = kh0;
                  if (dy == 2 && dx == 0) corner = k0w;
                  if (dy == 2 && dx == 2) corner = k00;
                  Sv[dy * 3 + dx] = T - rowx - colx + corner;
                }
            }
            if (!fixed_stats) grp.sync();  // the partial sums are dead: S takes their place
            if (et < 64) {
#pragma unroll
              for (int k = 0; k < 9; ++k) S[k * 64 + et] = Sv[k];
            }
            grp.sync();
            if (q == 0 && egrp == 0) DFIR_TRACE(8, 44);
            {
              // 2 threads per output channel split the taps (warps 0-1: even taps, warps 2-3: odd taps, so that a warp reads
              // one S chunk - a pure broadcast - and 32 consecutive weight rows - conflict-free through the 128B swizzle: the
              // phase is bound by shared-memory wavefronts, 72 KB of weights per image); eight independent chains (one per
              // position inside a 16-byte weight chunk), summed in a fixed order - the result depends on the layer only,
              // never on the band or the CTA
              const int co = et & 63, part2 = et >> 6;
              float ac[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
              const float4* S4 = reinterpret_cast<const float4*>(S);
              for (int tap = part2; tap < 9; tap += 2) {
                const uint8_t* wr = wsm + (tap * 64 + co) * 128;
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                  const uint4 raw = *reinterpret_cast<const uint4*>(wr + ((ch ^ (co & 7)) << 4));
                  const float4 sa = S4[tap * 16 + ch * 2], sb = S4[tap * 16 + ch * 2 + 1];
                  ac[0] = fmaf(__uint_as_float(raw.x << 16), sa.x, ac[0]);
                  ac[1] = fmaf(__uint_as_float(raw.x & 0xffff0000u), sa.y, ac[1]);
                  ac[2] = fmaf(__uint_as_float(raw.y << 16), sa.z, ac[2]);
                  ac[3] = fmaf(__uint_as_float(raw.y & 0xffff0000u), sa.w, ac[3]);
                  ac[4] = fmaf(__uint_as_float(raw.z << 16), sb.x, ac[4]);
                  ac[5] = fmaf(__uint_as_float(raw.z & 0xffff0000u), sb.y, ac[5]);
                  ac[6] = fmaf(__uint_as_float(raw.w << 16), sb.z, ac[6]);
                  ac[7] = fmaf(__uint_as_float(raw.w & 0xffff0000u), sb.w, ac[7]);
                }
              }
              const float acc0 = (ac[0] + ac[2]) + (ac[4] + ac[6]), acc1 = (ac[1] + ac[3]) + (ac[5] + ac[7]);
              tmp[640 + et] = acc0 + acc1;
            }
            grp.sync();
            if (q == 0 && egrp == 0) DFIR_TRACE(8, 45);
            if (et < 64) {
              y_s[et] = bias_s[et] + (tmp[640 + et] + tmp[640 + 64 + et]) * inv_hw;
              // training forward: the backward needs the pooled mean (every CTA touching image b writes the same bits)
              if (a.ymean_out != nullptr) a.ymean_out[static_cast<size_t>(ba) * 64 + et] = y_s[et];
            }
            grp.sync();
            if (q == 0 && egrp == 0) DFIR_TRACE(8, 46);
            attn_vector(grp, a.ca_style, cap, 64, a.ca_R, a.ca_M, attr_s, y_s, s_s, tmp);
            if (q == 0 && egrp == 0) DFIR_TRACE(8, 47);
            if (et < 64) svec_s[(b - bimg_first) * 64 + et] = s_s[et] * sqv;
            grp.sync();
            if (q == 0 && egrp == 0) DFIR_TRACE(8, 48);
          }
        }
      } else {
        grid_dep_wait();
        if constexpr (EPI == EPI_RELU_STATS || EPI == EPI_RELU_STATS_W) {
          // the statistics buffer the NEXT block's conv1 accumulates into: its last reader (the previous conv2) is complete
          if (a.istats_clear != nullptr) {
            const int nth = 128 * kEpiGroups;
            for (int i = blockIdx.x * nth + egrp * 128 + et; i < a.B * 576; i += gridDim.x * nth) a.istats_clear[i] = 0;
          }
        }
      }
      if constexpr (kTwoEpi && kScaleSkip) {
        if (a.epi_stats) named_bar_sync(6, 256);  // the attention vectors (svec_s) of both groups are complete
        if (q == 0 && egrp == 0) DFIR_TRACE(8, 49);
      }
      // (col, y, b, seg) of output row g, advanced without integer divisions (they cost ~130 clk each per row)
      int col = (g0 + egrp) / H, y = (g0 + egrp) % H;
      int b = col / nseg, seg = col % nseg;
      auto next_row = [&]() {
        y += kEpiGroups;
        while (y >= H) {
          y -= H;
          ++col;
          if (++seg == nseg) {
            seg = 0;
            ++b;
          }
        }
      };
      for (int g = g0 + egrp, it = egrp; g < g1; g += kEpiGroups, it += kEpiGroups, next_row()) {
        const int x = seg * 128 + m;
        const bool valid = x < a.W;
        const int acc = it % kAcc;
        if (probe) g_dfir_progress[8 + q] = it + 1;
        // EPI_SCALE_SKIP: the fp32 skip row travels global -> smem with cp.async in two halves of 32 channels (8 x 16 B
        // per thread and half, coalesced, no registers held).  Each thread later reads back exactly the 16-byte slots
        // it copied itself, so no barrier is needed, only cp.async.wait_group.
        const bool has_skip = a.skip_f32 != nullptr;  // dgrad launches without a skip: out = acc * s + b * s
        // Both 32-channel halves of a row have their own buffer, so the load of a half is issued a full pass (tile
        // write + barriers + global pass) before it is consumed: the copy of half h of the next row goes out as soon as
        // half h of this row has been read.  One commit group per call, also when nothing is left to load, so that
        // cp.async.wait_group 1 always means "the older of the two outstanding halves has landed".
        float* skipbuf_g = skipbuf + egrp * (128 * 32);
        auto issue_skip = [&](int gg, int hh) {
          if (!has_skip) return;
          if (gg >= g1) {
            asm volatile("cp.async.commit_group;" ::: "memory");
            return;
          }
          float* sdst = skipbuf_g + hh * (2 * 128 * 32);
          const int colg = gg / H;
          const int segg = colg % nseg;
          const int npxg = min(128, a.W - segg * 128);
          const float* src = a.skip_f32 + ((static_cast<size_t>(colg / nseg) * a.H + (gg % H)) * a.W + segg * 128) * 64 + hh * 32;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int idx = i * 128 + et;  // pixel idx >> 3, 16-byte chunk idx & 7
            if ((idx >> 3) < npxg && !exp_no_skipld) {
              if (a.use_hints)
                asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_u32(sdst + idx * 4)),
                             "l"(src + static_cast<size_t>(idx >> 3) * 64 + (idx & 7) * 4), "l"(a.pol_skip)
                             : "memory");
              else
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sdst + idx * 4)),
                             "l"(src + static_cast<size_t>(idx >> 3) * 64 + (idx & 7) * 4)
                             : "memory");
            }
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
        };
        if constexpr (EPI == EPI_SCALE_SKIP) {
          if (it == egrp) {
            issue_skip(g, 0);
            issue_skip(g, 1);
          }
          if (has_skip && !exp_no_pf && g + 2 * kEpiGroups < g1) {  // pull the skip row this group needs two rows from now into L2
            const int g2 = g + 2 * kEpiGroups, col2 = g2 / H;
            const size_t e2 = ((static_cast<size_t>(col2 / nseg) * a.H + (g2 % H)) * a.W + (col2 % nseg) * 128) * 64;
            const int npx2 = min(128, a.W - (col2 % nseg) * 128);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const int line = k * 128 + et;  // 256 lines of 128 B = one 128-px fp32 row
              if (line * 32 < npx2 * 64)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a.skip_f32 + e2 + static_cast<size_t>(line) * 32));
            }
          }
        }
        if (q == 0 && !(tile_probe && egrp == 1)) DFIR_TRACE(8 + 3 * egrp, it >> (kTwoEpi ? 1 : 0));  // epilogue: waiting for the accumulator
        mbar_wait(&tfull[acc], (it / kAcc) & 1, 8);
        if (q == 0 && !(tile_probe && egrp == 1)) DFIR_TRACE(9 + 3 * egrp, it >> (kTwoEpi ? 1 : 0));  // epilogue: accumulator complete
        tcgen05_fence_after();
        if (exp_no_epi) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&go[acc]);
          continue;
        }
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * NT;

        if constexpr (EPI == EPI_TAIL_NCHW) {
          uint32_t r0[16];
          tmem_ld_32x32b_x16(taddr, r0);
          tmem_ld_wait();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&go[acc]);
          if (valid) {  // fp32 NCHW output, a.cout real channels (<= NT)
            const size_t plane = static_cast<size_t>(a.H) * a.W;
            float* o = a.out_f32 + (static_cast<size_t>(b) * a.cout) * plane + static_cast<size_t>(y) * a.W + x;
#pragma unroll
            for (int c = 0; c < NT; ++c)
              if (c < a.cout) o[c * plane] = __uint_as_float(r0[c]) + bias_s[c] + (a.tail_accumulate ? o[c * plane] : 0.f);
          }
        } else if constexpr (kHL) {
          uint32_t ra[32], rb[32];  // 16x256b fragments: pixels pr, pr + 8 (ra) and pr + 16, pr + 24 (rb), 16 channels each
          tmem_ld_16x256b_x8(taddr, ra);
          tmem_ld_16x256b_x8(taddr + (16u << 16), rb);
          tmem_ld_wait();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&go[acc]);  // accumulator back to the MMA thread
          if (q == 0 && !(tile_probe && egrp == 1)) DFIR_TRACE(10 + 3 * egrp, it >> 1);
          const int pr = lane >> 2, cq = lane & 3;
          if (b != cur_img) {  // (uniform) new image: this thread's 16 scale / bias * scale values
#pragma unroll
            for (int n = 0; n < 8; ++n)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int c = 8 * n + 2 * cq + e;
                const float sc = a.epi_stats ? svec_s[(b - bimg_first) * 64 + c]
                                             : (a.svec != nullptr ? a.svec[static_cast<size_t>(flip ? a.B - 1 - b : b) * 64 + c] : 1.f);
                hl_s[2 * n + e] = sc;
                hl_bs[2 * n + e] = bias_s[c] * sc;
              }
            cur_img = b;
          }
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            // 16-pixel tiles (bf16 lo): a tile per half; 32-pixel tiles (8-bit lo): one tile per row, entered at half 0
            const bool tile_start = kTilePx == 16 || half == 0, tile_end = kTilePx == 16 || half == 1;
            uint8_t* buf = hl_base + hl_slot * (8 * kHI);
            const bool late_issue = !kLo8 && (a.debug_probe & 131072) != 0;  // A/B switch (DFIR_DEBUG_PROBE)
            if (tile_start) {
              if (lane == 0 && !late_issue) {
                // the previous tile must have left its buffer (its store has read it), which takes the tile after next
                // (three buffers of 16 pixels) / the next row's tile (two buffers of 32 pixels)
                tma_store_wait_read<0>();
                hl_issue();
              }
              if (tile_probe && q == 0 && egrp == 0 && half == 0) DFIR_TRACE(11, it >> 1);  // next load issued
              mbar_wait(&hl_bar[hl_slot], hl_phase, 10);
              if (tile_probe && q == 0 && egrp == 0 && half == 0) DFIR_TRACE(12, it >> 1);  // tile landed
            }
            // word (pixel p, channels 8 n + 2 cq + {0,1}) of a tile: p * 128 + ((n ^ (p & 7)) << 4) + 4 cq (TMA 128B swizzle)
            uint8_t* wbase = buf + pr * 128 + 4 * cq + (kTilePx == 32 ? half * 2048 : 0);
            const uint32_t sbuf = hl_sbase + hl_slot * (8 * kHI);   // the buffer as a shared-window address
#pragma unroll
            for (int sl = 0; sl < 2; ++sl) {
              // all 16 loads of a pixel first, then the arithmetic, then the 16 stores: written as load / update / store
              // per word the compiler must keep every load behind the previous word's store (same buffer), which
              // serialises 16 shared-memory round trips (measured: 3000 clk per tile instead of ~700)
              uint32_t hw[8], lw[8];
              if constexpr (kLo8) {
                // 8-bit lo plane: the stream value is a 24-BIT float X (sign, exponent, 15 mantissa bits = 16 significant bits);
                // hi = the bf16 nearest to X (ties away from zero), q = (X - hi) in units of X's last bit, taken on the BIT
                // PATTERNS: bits(X) = (hi << 16) + (q << 8), q in [-128, 127] - exact across binade boundaries, no exponent
                // arithmetic, no conversions.  The lo plane is private to these kernels, so its 64 bytes per pixel are stored in
                // the order the accumulator fragment wants: byte cq * 16 + 2 n + e holds channel 8 n + 2 cq + e, i.e. the 16
                // channels of a thread are ONE 16-byte word (one 128-bit shared-memory access per pixel instead of eight 16-bit
                // ones).  Tile of 32 px x 64 B, TMA SWIZZLE_64B: pixel p, 16-byte chunk k at p * 64 + ((k ^ ((p >> 1) & 3)) << 4).
                const int p = pr + 8 * sl + (kTilePx == 32 ? 16 * half : 0);
                const uint32_t lo_a = sbuf + kHlHiBytes + p * 64 + ((cq ^ ((p >> 1) & 3)) << 4);
                // hi word n of the pixel: the buffer is 1024-byte aligned, so + ((n ^ pr) << 4) is ^ (pr << 4) ^ (n << 4)
                const uint32_t hi_a = (sbuf + pr * 128 + 4 * cq + half * 2048 + sl * 1024) ^ (pr << 4);
                uint32_t lq[4] = {0u, 0u, 0u, 0u};
                if (!exp_no_lo) {
                  const uint4 t4 = lds_v4(lo_a);
                  lq[0] = t4.x; lq[1] = t4.y; lq[2] = t4.z; lq[3] = t4.w;
                }
#pragma unroll
                for (int n = 0; n < 8; ++n) hw[n] = exp_no_hi ? 0x3f803f80u : lds_u32(hi_a ^ (n << 4));
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                  const uint32_t* src = half ? rb : ra;
                  // (sign-extended q) << 8 by one byte permute each: bytes [0, q, sign, sign]; q pair n = half (n & 1) of lq[n >> 1]
                  const uint32_t qw = lq[n >> 1];
                  const float x0 = __uint_as_float((hw[n] << 16) + prmt(qw, 0u, (n & 1) ? 0xAA24u : 0x8804u));
                  const float x1 = __uint_as_float((hw[n] & 0xffff0000u) + prmt(qw, 0u, (n & 1) ? 0xBB34u : 0x9914u));
                  float o0 = fmaf(__uint_as_float(src[4 * n + 2 * sl]), hl_s[2 * n], hl_bs[2 * n]) + x0;
                  float o1 = fmaf(__uint_as_float(src[4 * n + 2 * sl + 1]), hl_s[2 * n + 1], hl_bs[2 * n + 1]) + x1;
                  if (a.relu_out) {  // (last K-chunk of a wide conv + ReLU)
                    o0 = fmaxf(o0, 0.f);
                    o1 = fmaxf(o1, 0.f);
                  }
                  // round to 24 bits (+0x80) and to the nearest bf16 (+0x8000) in one add: u = bits(o) + 0x8080.  hi = the top
                  // half of u; q = byte 1 of (u & 0xffff) - 0x8000 = byte 1 of u with its top bit flipped (v - 128 = v ^ 0x80 mod 256)
                  const uint32_t u0 = __float_as_uint(o0) + 0x8080u, u1 = __float_as_uint(o1) + 0x8080u;
                  hw[n] = prmt(u0, u1, 0x7632u);
                  lw[n] = prmt(u0, u1, 0x0051u);   // (low 16 bits: the q pair before the sign flip)
                }
#pragma unroll
                for (int n = 0; n < 8; ++n)
                  if (!exp_no_hi) sts_u32(hi_a ^ (n << 4), hw[n]);
                if (!exp_no_lo) {
                  uint4 o4;
                  o4.x = prmt(lw[0], lw[1], 0x5410u) ^ 0x80808080u;
                  o4.y = prmt(lw[2], lw[3], 0x5410u) ^ 0x80808080u;
                  o4.z = prmt(lw[4], lw[5], 0x5410u) ^ 0x80808080u;
                  o4.w = prmt(lw[6], lw[7], 0x5410u) ^ 0x80808080u;
                  sts_v4(lo_a, o4);
                }
              } else {
#pragma unroll
              for (int n = 0; n < 8; ++n) {
                hw[n] = *reinterpret_cast<const uint32_t*>(wbase + sl * 1024 + ((n ^ pr) << 4));
                lw[n] = *reinterpret_cast<const uint32_t*>(wbase + 2048 + sl * 1024 + ((n ^ pr) << 4));
              }
#pragma unroll
              for (int n = 0; n < 8; ++n) {
                const uint32_t* src = half ? rb : ra;
                const float x0 = __uint_as_float(hw[n] << 16) + __uint_as_float(lw[n] << 16);
                const float x1 = __uint_as_float(hw[n] & 0xffff0000u) + __uint_as_float(lw[n] & 0xffff0000u);
                const float o0 = fmaf(__uint_as_float(src[4 * n + 2 * sl]), hl_s[2 * n], hl_bs[2 * n]) + x0;
                const float o1 = fmaf(__uint_as_float(src[4 * n + 2 * sl + 1]), hl_s[2 * n + 1], hl_bs[2 * n + 1]) + x1;
                const uint32_t nh = pack_bf16x2(o0, o1);
                hw[n] = nh;
                lw[n] = pack_bf16x2(o0 - __uint_as_float(nh << 16), o1 - __uint_as_float(nh & 0xffff0000u));
              }
#pragma unroll
              for (int n = 0; n < 8; ++n) {
                *reinterpret_cast<uint32_t*>(wbase + sl * 1024 + ((n ^ pr) << 4)) = hw[n];
                *reinterpret_cast<uint32_t*>(wbase + 2048 + sl * 1024 + ((n ^ pr) << 4)) = lw[n];
              }
              }
            }
            if (tile_probe && q == 0 && egrp == 0 && half == 0) DFIR_TRACE(13, it >> 1);  // tile (half) updated
            if (!tile_end) continue;
            fence_proxy_async_smem();
            __syncwarp();
            if (tile_probe && q == 0 && egrp == 0 && half == 1) DFIR_TRACE(15, it >> 1);  // last tile of the row fenced
            if (lane == 0) {
              const int xs = seg * 128 + q * 32 + (kTilePx == 16 ? half * 16 : 0);
              const int ya = flip ? H - 1 - y : y, ba = flip ? a.B - 1 - b : b;
              if (a.use_hints) {
                tma_store_4d_hint(&hl.m[2], buf, 0, xs, ya, ba, a.pol_out);
                if (a.hl_store_lo) tma_store_4d_hint(&hl.m[3], buf + kHlHiBytes, 0, xs, ya, ba, a.pol_f32);
              } else {
                tma_store_4d(&hl.m[2], buf, 0, xs, ya, ba);
                if (a.hl_store_lo) tma_store_4d(&hl.m[3], buf + kHlHiBytes, 0, xs, ya, ba);
              }
              tma_store_commit();
              if (late_issue) {
                tma_store_wait_read<1>();
                hl_issue();
              }
            }
            if (++hl_slot == kHS) {
              hl_slot = 0;
              hl_phase ^= 1;
            }
          }
          if (q == 0 && !(tile_probe && egrp == 1)) DFIR_TRACE(14 + egrp, it >> 1);
        } else if constexpr (EPI == EPI_RELU_STATS_W) {
          // ---- conv1 + ReLU + image statistics, WARP-AUTONOMOUS (no block-level barrier): every epilogue warp stages its own
          // 32 pixels (4 KB, two private buffers) and stores them with its own TMA operation; the statistics of
          // pool-by-linearity are accumulated per THREAD in 32-bit fixed point (2^-12 units, integer addition: independent of
          // how rows are grouped into bands, hence still batch-invariant) and reduced over the warp only when the image
          // changes.  Rows y = 0 / H - 1 (first / last row sums, corners) add their parts with atomics directly.
          uint32_t ra[32], rb[32];  // pixels pr, pr + 8 (ra) and pr + 16, pr + 24 (rb) of this warp's quarter
          tmem_ld_16x256b_x8(taddr, ra);
          tmem_ld_16x256b_x8(taddr + (16u << 16), rb);
          tmem_ld_wait();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&go[acc]);
          if (q == 0) DFIR_TRACE(10 + 3 * egrp, it >> 1);
          const int pr = lane >> 2, cq = lane & 3;
          auto flush_image = [&](int img) {  // this thread's accumulators -> the image's 64-bit sums (2^-24 units)
            // transposing butterfly over the 8 pixel rows of the fragment (lane bits 2..4), on integers: on exit w_t[0], w_t[1]
            // are the warp's sums of channels 8 n + 2 cq + {0, 1}, n = lane bits (4, 3, 2)
#pragma unroll
            for (int step = 0; step < 3; ++step) {
              const int nv = 16 >> step;
              const int mask = 16 >> step;
              const bool upper = (lane & mask) != 0;
#pragma unroll
              for (int i = 0; i < nv / 2; ++i) {
                const int send = upper ? w_t[i] : w_t[i + nv / 2];
                const int keep = upper ? w_t[i + nv / 2] : w_t[i];
                w_t[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
              }
            }
            const int nblk = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
            unsigned long long* dst = reinterpret_cast<unsigned long long*>(a.istats) + static_cast<size_t>(img) * 576;
            atomicAdd(dst + 8 * nblk + 2 * cq, static_cast<unsigned long long>(static_cast<long long>(w_t[0]) << 12));
            atomicAdd(dst + 8 * nblk + 2 * cq + 1, static_cast<unsigned long long>(static_cast<long long>(w_t[1]) << 12));
            if (w_cw != 0) {  // this lane owns the first (1) or last (2) column of the image's rows it saw
              unsigned long long* cdst = dst + (w_cw == 1 ? 64 : 128);
#pragma unroll
              for (int j = 0; j < 16; ++j)
                atomicAdd(cdst + 8 * (j >> 1) + 2 * cq + (j & 1), static_cast<unsigned long long>(static_cast<long long>(w_c[j]) << 12));
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) w_t[j] = w_c[j] = 0;
            w_cw = 0;
          };
          if (b != cur_img) {  // (uniform) image change
            if (cur_img >= 0) flush_image(cur_img);
            cur_img = b;
          }
          float bias_r[16];
#pragma unroll
          for (int n = 0; n < 8; ++n) {
            bias_r[2 * n] = bias_s[8 * n + 2 * cq];
            bias_r[2 * n + 1] = bias_s[8 * n + 2 * cq + 1];
          }
          uint8_t* wbuf = stage + (egrp * 4 + q) * 8192 + ((it >> 1) & 1) * 4096;   // this warp's staging buffer of this row
          if (lane == 0) tma_store_wait_read<1>();   // the store that used this buffer two rows ago has read it
          __syncwarp();
          float sums[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) sums[j] = 0.f;
          const int x0 = seg * 128 + q * 32 + pr;
          const bool edge_row = y == 0 || y == a.H - 1;
          unsigned long long* irow = reinterpret_cast<unsigned long long*>(a.istats) + static_cast<size_t>(b) * 576;
          uint8_t* strow = wbuf + pr * 128 + 4 * cq;   // (pixel & 7) == pr for all four pixels of the thread
#pragma unroll
          for (int sl = 0; sl < 4; ++sl) {   // pixel slots pr, pr + 8, pr + 16, pr + 24
            const int xs = x0 + 8 * sl;
            const bool ok = xs < a.W;
            float v[16];
#pragma unroll
            for (int n = 0; n < 8; ++n) {
              const uint32_t* src = sl < 2 ? ra : rb;
              v[2 * n] = fmaxf(__uint_as_float(src[4 * n + 2 * (sl & 1)]) + bias_r[2 * n], 0.f);
              v[2 * n + 1] = fmaxf(__uint_as_float(src[4 * n + 2 * (sl & 1) + 1]) + bias_r[2 * n + 1], 0.f);
            }
#pragma unroll
            for (int n = 0; n < 8; ++n)
              *reinterpret_cast<uint32_t*>(strow + sl * 1024 + ((n ^ pr) << 4)) = pack_bf16x2(v[2 * n], v[2 * n + 1]);
            if (ok) {
#pragma unroll
              for (int j = 0; j < 16; ++j) sums[j] += v[j];
              if (xs == 0 || xs == a.W - 1) {  // first / last column: this lane owns 16 channels of that pixel and keeps their
                w_cw = xs == 0 ? 1 : 2;        // sums in registers (a thread never owns both columns: host check)
#pragma unroll
                for (int j = 0; j < 16; ++j) w_c[j] += __float2int_rn(v[j] * 4096.f);
              }
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&hl.m[0], wbuf, 0, seg * 128 + q * 32, y, b);
            tma_store_commit();
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) w_t[j] += __float2int_rn(sums[j] * 4096.f);
          if (edge_row) {
            // (uniform, two rows per image) first / last row sums R0 / RL and the corner pixels K00, K0W, KH0, KHW: straight to
            // the image's sums.  Kept out of the per-row path: ~200 conditional atomics in the unrolled loops cost every row
            // ~1000 instructions of predicates and branches (measured: conv1 36 -> 59 us).
            for (int e = 0; e < 2; ++e) {
              if ((e == 0 && y != 0) || (e == 1 && y != a.H - 1)) continue;
              unsigned long long* rdst = irow + (e == 0 ? 192 : 256);
#pragma unroll
              for (int j = 0; j < 16; ++j)
                atomicAdd(rdst + 8 * (j >> 1) + 2 * cq + (j & 1),
                          static_cast<unsigned long long>(static_cast<long long>(__float2int_rn(sums[j] * 4096.f)) << 12));
#pragma unroll
              for (int sl = 0; sl < 4; ++sl) {   // (static indices into the fragment registers: no local-memory copy)
                const int xs = x0 + 8 * sl;
                if (xs >= a.W || (xs != 0 && xs != a.W - 1)) continue;
                const bool kfirst = xs == 0, klast = xs == a.W - 1;   // (W == 1: both)
                unsigned long long* kdst = irow + 320 + 128 * e;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const uint32_t* src = sl < 2 ? ra : rb;
                  const float val = fmaxf(__uint_as_float(src[4 * (j >> 1) + 2 * (sl & 1) + (j & 1)]) + bias_r[j], 0.f);
                  const unsigned long long fv =
                      static_cast<unsigned long long>(static_cast<long long>(__float2int_rn(val * 4096.f)) << 12);
                  const int ch = 8 * (j >> 1) + 2 * cq + (j & 1);
                  if (kfirst) atomicAdd(kdst + ch, fv);
                  if (klast) atomicAdd(kdst + 64 + ch, fv);
                }
              }
            }
          }
          if (g + kEpiGroups >= g1 && cur_img >= 0) {  // last row of this warp: flush
            flush_image(cur_img);
            cur_img = -1;
          }
          if (q == 0) DFIR_TRACE(14 + egrp, it >> 1);
        } else if constexpr (EPI == EPI_SCALE_SKIP) {
          // Per half of 32 channels: v = acc * s + bias * s (or r = acc + bias when the training forward saves r) into
          // an fp32 tile in smem (chunk-rotated: conflict free for the pixel-major writes and the coalesced reads),
          // then a coalesced pass adds the fp32 skip and writes the fp32 stream + its bf16 copy to global memory.
          float* tile = reinterpret_cast<float*>(stage) + egrp * (128 * 32);  // [128 px][32] fp32 = 16 KB per group
          float* sc_s = pool_s + egrp * 192;              // [64] scale, [64] bias*scale, [64] real scale
          const bool save_r = a.r_out != nullptr;         // training forward: tile = acc + b, scale applied below
          const uint32_t bar_a = 1 + 2 * egrp, bar_b = 2 + 2 * egrp;
          const int npx = min(128, a.W - seg * 128);      // valid pixels of this row segment
          const int c4 = et & 7, pq = et >> 3;            // this thread: 4-channel group c4 of pixels pq + 16 i
          // Both 32-column halves of the accumulator are read early (the second one right after the first has been
          // written to the tile), so the MMA warp gets the accumulator back after ~2 TMEM reads instead of after the
          // first half's global-memory pass: with two accumulators the hold time, not the MMA rate, set the pace
          // (measured: 58 us per 32-image launch with every global access of the epilogue removed, 38 us without the
          // epilogue).
          uint32_t rv[32];
          auto write_tile = [&](int h) {
            if (exp_no_tile) return;
            // plain copy: scale and bias are applied in the pass, where each thread owns one 4-channel group
            float4* trow = reinterpret_cast<float4*>(tile + m * 32);
#pragma unroll
            for (int c = 0; c < 8; ++c)
              trow[(c + m) & 7] = make_float4(__uint_as_float(rv[4 * c + 0]), __uint_as_float(rv[4 * c + 1]),
                                              __uint_as_float(rv[4 * c + 2]), __uint_as_float(rv[4 * c + 3]));
          };
          // `fast` (compile-time tag): the inference chain's case — full 128-px row, fp32 skip and fp32 output present,
          // no saved r, no ReLU, no cache hints — as straight-line code; everything else takes the general form.  The
          // epilogue loop is instruction-fetch sensitive (ncu: 11 % of its samples stall on no_inst).
          const bool fast_pass = !save_r && has_skip && !a.relu_out && !a.use_hints && a.out_f32 != nullptr && npx == 128 &&
                                 !exp_no_f32st && !exp_no_bfst && (a.debug_probe & 16384) == 0;  // bit 16384: A/B switch
          auto pass_impl = [&](int h, auto fast) {
            constexpr bool F = decltype(fast)::value;
            if (exp_no_tile || exp_no_pass) return;
            const size_t e0 = ((static_cast<size_t>(b) * a.H + y) * a.W + seg * 128 + pq) * 64 + h * 32 + c4 * 4;
            float* o32 = a.out_f32 != nullptr ? a.out_f32 + e0 : nullptr;
            __nv_bfloat16* obf = a.out_bf16_direct + e0;
            __nv_bfloat16* rbf = save_r ? a.r_out + e0 : nullptr;
            const float4* t4 = reinterpret_cast<const float4*>(tile);
            const float4* sk4 = reinterpret_cast<const float4*>(skipbuf_g + h * (2 * 128 * 32)) + et;
            const float4 s4 = reinterpret_cast<const float4*>(sc_s)[h * 8 + c4];
            const float4 b4 = reinterpret_cast<const float4*>(sc_s + 64)[h * 8 + c4];
            const float4 sr4 = reinterpret_cast<const float4*>(sc_s + 128)[h * 8 + c4];
            asm volatile("cp.async.wait_group 1;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int p = i * 16 + pq;
              if (F || p < npx) {
                const float4 tv = t4[p * 8 + ((c4 + p) & 7)];
                float4 o;
                o.x = fmaf(tv.x, s4.x, b4.x); o.y = fmaf(tv.y, s4.y, b4.y);
                o.z = fmaf(tv.z, s4.z, b4.z); o.w = fmaf(tv.w, s4.w, b4.w);
                if (!F && save_r) {  // r = conv + b goes out as bf16 for the backward; the stream gets r * s + skip
                  uint2 rk;
                  rk.x = pack_bf16x2(o.x, o.y);
                  rk.y = pack_bf16x2(o.z, o.w);
                  *reinterpret_cast<uint2*>(rbf + i * 1024) = rk;
                  o.x *= sr4.x; o.y *= sr4.y; o.z *= sr4.z; o.w *= sr4.w;
                }
                if (F || has_skip) {
                  const float4 sk = sk4[i * 128];
                  o.x += sk.x; o.y += sk.y; o.z += sk.z; o.w += sk.w;
                }
                if (!F && a.relu_out) {
                  o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                }
                uint2 pk;
                pk.x = pack_bf16x2(o.x, o.y);
                pk.y = pack_bf16x2(o.z, o.w);
                if (F) {
                  *reinterpret_cast<float4*>(o32 + i * 1024) = o;
                  *reinterpret_cast<uint2*>(obf + i * 1024) = pk;
                } else if (a.use_hints) {
                  if (o32 != nullptr) st_global_v4_hint(o32 + i * 1024, o, a.pol_f32);
                  st_global_v2_hint(obf + i * 1024, pk, a.pol_out);
                } else {
                  if (o32 != nullptr && !exp_no_f32st) *reinterpret_cast<float4*>(o32 + i * 1024) = o;
                  if (!exp_no_bfst) *reinterpret_cast<uint2*>(obf + i * 1024) = pk;
                }
              }
            }
          };
          auto pass = [&](int h) {
            if (fast_pass) pass_impl(h, std::true_type{});
            else pass_impl(h, std::false_type{});
          };
          named_bar_sync(bar_a, 128);  // the previous row's second pass has finished with the tile (and with sc_s)
          if (b != cur_img) {          // (uniform) new image: stage its scale vector
            if (et < 64) {
              const float sc = a.epi_stats ? svec_s[(b - bimg_first) * 64 + et]
                                           : (a.svec != nullptr ? a.svec[static_cast<size_t>(b) * 64 + et] : 1.f);
              sc_s[et] = save_r ? 1.f : sc;
              sc_s[64 + et] = save_r ? bias_s[et] : bias_s[et] * sc;
              sc_s[128 + et] = sc;
            }
            cur_img = b;
            named_bar_sync(bar_b, 128);
          }
          tmem_ld_32x32b_x32(taddr, rv);
          tmem_ld_wait();
          write_tile(0);
          tmem_ld_32x32b_x32(taddr + 32, rv);
          tmem_ld_wait();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&go[acc]);  // accumulator back to the MMA warp
          if (q == 0) DFIR_TRACE(10 + 3 * egrp, it >> 1);
          named_bar_sync(bar_b, 128);
          pass(0);
          issue_skip(g + kEpiGroups, 0);  // this thread's half-0 slots are free again: next row of this group
          named_bar_sync(bar_a, 128);
          write_tile(1);
          named_bar_sync(bar_b, 128);
          pass(1);
          issue_skip(g + kEpiGroups, 1);
          if (q == 0) DFIR_TRACE(14 + egrp, it >> 1);
        } else {
          // one group: the two staging buffers alternate.  Two groups: each group alternates between two buffers of its
          // own (the skip buffers of the scale+skip epilogue are free here), so a row never waits for the TMA store of the
          // group's previous row to drain its staging tile (measured: ~800 clk per row on the epilogue's critical path)
          const int sb = kTwoEpi ? (egrp * 2 + ((it >> 1) & 1)) : (it & 1);
          uint8_t* st = stage + sb * kStageBytes;
          float* pool_g = pool_s + egrp * 384;
          const uint32_t bar_c = 1 + 2 * egrp, bar_d = 2 + 2 * egrp;
          // Inference epilogues (bias / ReLU / per-row channel sums): the accumulator is read in the 16x256b fragment
          // layout — a thread owns 4 pixels x 16 channels instead of 1 pixel x 64 channels — so the per-row channel sums
          // need 3 butterfly levels over 16 values (14 shuffles) after 48 in-register adds instead of 5 levels over 2 x 32
          // values (62 shuffles), and the bias lives in 16 registers for the whole kernel.
          constexpr bool kQuad = EPI == EPI_BIAS || EPI == EPI_BIAS_RELU || EPI == EPI_BIAS_POOL || EPI == EPI_RELU_STATS;
          if constexpr (kQuad) {
            uint32_t ra[32], rb[32];  // pixels pr, pr + 8 (ra) and pr + 16, pr + 24 (rb) of this warp's quarter
            tmem_ld_16x256b_x8(taddr, ra);
            tmem_ld_16x256b_x8(taddr + (16u << 16), rb);
            tmem_ld_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&go[acc]);
            if (q == 0) DFIR_TRACE(10 + 3 * egrp, it >> (kTwoEpi ? 1 : 0));
            if (et == 0) {  // the TMA store that read this staging buffer has drained it
              tma_store_wait_read<1>();
            }
            named_bar_sync(bar_c, 128);
            if (q == 0 && egrp == 0) DFIR_TRACE(11, it >> 1);  // staging tile free
            const int pr = lane >> 2, cq = lane & 3;       // pixel row of the fragment, channel pair within a block of 8
            float bias_r[16];
#pragma unroll
            for (int n = 0; n < 8; ++n) {
              bias_r[2 * n] = bias_s[8 * n + 2 * cq];
              bias_r[2 * n + 1] = bias_s[8 * n + 2 * cq + 1];
            }
            float sums[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) sums[j] = 0.f;
            const int x0 = seg * 128 + q * 32 + pr;         // image column of this thread's first pixel
            // staging word (pixel slot sl, channel block n) by shared-window address: the tile is 1024-byte aligned, so
            // + ((n ^ pr) << 4) is ^ (pr << 4) ^ (n << 4); (pixel & 7) == pr for all four pixels of the thread
            const uint32_t st_a = (smem_u32(st) + (q * 32 + pr) * 128 + 4 * cq) ^ (pr << 4);
#pragma unroll
            for (int sl = 0; sl < 4; ++sl) {                // pixel slots: pr, pr + 8, pr + 16, pr + 24
              const bool ok = x0 + 8 * sl < a.W;
              float v[16];
#pragma unroll
              for (int n = 0; n < 8; ++n) {
                const uint32_t* src = sl < 2 ? ra : rb;
                v[2 * n] = __uint_as_float(src[4 * n + 2 * (sl & 1)]) + bias_r[2 * n];
                v[2 * n + 1] = __uint_as_float(src[4 * n + 2 * (sl & 1) + 1]) + bias_r[2 * n + 1];
              }
              if constexpr (EPI == EPI_BIAS_RELU || EPI == EPI_RELU_STATS) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
              }
              if constexpr (EPI == EPI_RELU_STATS) {
                // statistics of t in fp32 (before the bf16 rounding of the store): the rounding noise averages out
                // over the image (relative effect on the pooled mean ~1e-5) and the fp32 value is what the
                // reference's own pooled mean is made of
                const int xs = x0 + 8 * sl;
                if (a.istats != nullptr) {  // first / last column of the row go to the 64 accumulating threads via smem
                  if (ok && xs == 0) {
#pragma unroll
                    for (int n = 0; n < 8; ++n)
                      *reinterpret_cast<float2*>(pool_g + 256 + 8 * n + 2 * cq) = make_float2(v[2 * n], v[2 * n + 1]);
                  }
                  if (ok && xs == a.W - 1) {
#pragma unroll
                    for (int n = 0; n < 8; ++n)
                      *reinterpret_cast<float2*>(pool_g + 320 + 8 * n + 2 * cq) = make_float2(v[2 * n], v[2 * n + 1]);
                  }
                } else if (ok && (xs == 0 || xs == a.W - 1)) {
                  const size_t ro = (static_cast<size_t>(b) * a.H + y) * 64 + 2 * cq;
                  if (xs == 0) {
#pragma unroll
                    for (int n = 0; n < 8; ++n)
                      *reinterpret_cast<float2*>(a.col_first + ro + 8 * n) = make_float2(v[2 * n], v[2 * n + 1]);
                  }
                  if (xs == a.W - 1) {
#pragma unroll
                    for (int n = 0; n < 8; ++n)
                      *reinterpret_cast<float2*>(a.col_last + ro + 8 * n) = make_float2(v[2 * n], v[2 * n + 1]);
                  }
                }
              }
              if (!exp_skip_store) {
#pragma unroll
                for (int n = 0; n < 8; ++n) sts_u32((st_a + sl * 1024) ^ (n << 4), pack_bf16x2(v[2 * n], v[2 * n + 1]));
              }
              if constexpr (EPI == EPI_BIAS_POOL || EPI == EPI_RELU_STATS) {
                if (ok) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) sums[j] += v[j];
                }
              }
            }
            if (q == 0 && egrp == 0) DFIR_TRACE(12, it >> 1);  // values computed and staged
            if constexpr (EPI == EPI_BIAS_POOL || EPI == EPI_RELU_STATS) {
              // transposing butterfly over the 8 pixel rows of the fragment (lane bits 2..4): 8 + 4 + 2 shuffles; on exit
              // sums[0], sums[1] are the warp's 32-pixel sums of channels 8 n + 2 cq + {0, 1}, n = lane bits (4, 3, 2)
#pragma unroll
              for (int step = 0; step < 3; ++step) {
                const int nv = 16 >> step;
                const int mask = 16 >> step;
                const bool upper = (lane & mask) != 0;
#pragma unroll
                for (int i = 0; i < nv / 2; ++i) {
                  const float send = upper ? sums[i] : sums[i + nv / 2];
                  const float keep = upper ? sums[i + nv / 2] : sums[i];
                  sums[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
                }
              }
              const int nblk = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
              *reinterpret_cast<float2*>(pool_g + q * 64 + 8 * nblk + 2 * cq) = make_float2(sums[0], sums[1]);
            }
          } else {
          // The accumulator goes back to the MMA thread before anything else happens: a TMEM read of 32 columns costs
          // ~30 clk (tools/mma_probe.cu), everything after it (staging-buffer hand-over, math, stores) several hundred.
          uint32_t rv2[2][32];
          tmem_ld_32x32b_x32(taddr, rv2[0]);
          tmem_ld_32x32b_x32(taddr + 32, rv2[1]);
          tmem_ld_wait();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&go[acc]);
          if (q == 0) DFIR_TRACE(10 + 3 * egrp, it >> (kTwoEpi ? 1 : 0));
          if (et == 0) {  // the TMA store that read this staging buffer has drained it
            tma_store_wait_read<1>();
          }
          named_bar_sync(bar_c, 128);
          const size_t pix = (static_cast<size_t>(b) * a.H + y) * a.W + x;
#pragma unroll
          for (int h = 0; h < 2; ++h) {  // two halves of 32 channels bound the register footprint
            const uint32_t (&rv)[32] = rv2[h];
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(rv[i]) + bias_s[h * 32 + i];
            if constexpr (EPI == EPI_BIAS_RELU || EPI == EPI_RELU_STATS) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            if constexpr (EPI == EPI_RELU_STATS) {
              // statistics of t in fp32 (before the bf16 rounding of the store): the rounding noise averages out
              // over the image (relative effect on the pooled mean ~1e-5) and the fp32 value is what the
              // reference's own pooled mean is made of
              if (valid && (x == 0 || x == a.W - 1)) {
                float* dst = (x == 0 ? a.col_first : a.col_last) + (static_cast<size_t>(b) * a.H + y) * 64 + h * 32;
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                  *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                if (a.W == 1) {  // the only column is both first and last
                  float* d2 = a.col_last + (static_cast<size_t>(b) * a.H + y) * 64 + h * 32;
#pragma unroll
                  for (int i = 0; i < 32; i += 4)
                    *reinterpret_cast<float4*>(d2 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
              }
            }
            if constexpr (EPI == EPI_RELU_MASK) {
              // backward of ReLU: keep the gradient where the saved forward activation t is positive
              if (valid) {
                const uint4* mk = reinterpret_cast<const uint4*>(a.mask_bf16 + pix * 64 + h * 32);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  const uint4 raw = mk[c];
                  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float2 f = __bfloat1622float2(h2[j]);
                    if (!(f.x > 0.f)) v[8 * c + 2 * j] = 0.f;
                    if (!(f.y > 0.f)) v[8 * c + 2 * j + 1] = 0.f;
                  }
                }
              }
            }
            if constexpr (EPI == EPI_BIAS_SKIP) {
              if (valid) {
                const float4* sk = reinterpret_cast<const float4*>(a.skip_f32 + pix * 64 + h * 32);
                float4* o = reinterpret_cast<float4*>(a.out_f32 + pix * 64 + h * 32);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                  const float4 s = sk[c];
                  v[4 * c + 0] += s.x;
                  v[4 * c + 1] += s.y;
                  v[4 * c + 2] += s.z;
                  v[4 * c + 3] += s.w;
                  if (a.out_f32 != nullptr) o[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                }
              }
            }
            // bf16 -> swizzled staging row
            uint4* row = reinterpret_cast<uint4*>(st + m * 128);
            if (!exp_skip_store)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint4 pk;
              pk.x = pack_bf16x2(v[8 * c + 0], v[8 * c + 1]);
              pk.y = pack_bf16x2(v[8 * c + 2], v[8 * c + 3]);
              pk.z = pack_bf16x2(v[8 * c + 4], v[8 * c + 5]);
              pk.w = pack_bf16x2(v[8 * c + 6], v[8 * c + 7]);
              row[(h * 4 + c) ^ (m & 7)] = pk;
            }
            if constexpr (EPI == EPI_BIAS_POOL || EPI == EPI_RELU_STATS) {
              // per-row channel sums (EPI_BIAS_POOL: of the fp32 conv output; EPI_RELU_STATS: of rounded t)
              if (!valid) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.f;
              }
              warp_channel_sums32(v, lane);
              pool_g[q * 64 + h * 32 + lane] = v[0];
            }
          }
          }  // !kQuad
          if (q == 0 && egrp == 0) DFIR_TRACE(13, it >> 1);  // channel sums done
          fence_proxy_async_smem();
          named_bar_sync(bar_d, 128);
          if (q == 0 && egrp == 0) DFIR_TRACE(15, it >> 1);  // whole group past the second barrier
          if (et == 0 && !exp_skip_store) {
            if (a.use_hints) tma_store_4d_hint(&tmap_out, st, 0, seg * 128, y, b, a.pol_out);
            else tma_store_4d(&tmap_out, st, 0, seg * 128, y, b);
            tma_store_commit();
          }
          if constexpr (EPI == EPI_BIAS_POOL || EPI == EPI_RELU_STATS) {
            if (et < 64) {
              const float s = ((pool_g[et] + pool_g[64 + et]) + pool_g[128 + et]) + pool_g[192 + et];
              if (EPI == EPI_RELU_STATS && a.istats != nullptr) {
                if (b != cur_img) {  // (uniform) image change: flush the previous image's accumulators
                  if (cur_img >= 0) {
                    unsigned long long* dst = reinterpret_cast<unsigned long long*>(a.istats) + static_cast<size_t>(cur_img) * 576 + et;
                    atomicAdd(dst, static_cast<unsigned long long>(fx_t));
                    atomicAdd(dst + 64, static_cast<unsigned long long>(fx_c0));
                    atomicAdd(dst + 128, static_cast<unsigned long long>(fx_cl));
                  }
                  fx_t = fx_c0 = fx_cl = 0;
                  cur_img = b;
                }
                constexpr float kFx = 16777216.f;  // 2^24
                const long long fs = __float2ll_rn(s * kFx);
                fx_t += fs;
                unsigned long long* dst = reinterpret_cast<unsigned long long*>(a.istats) + static_cast<size_t>(b) * 576 + et;
                if (y == 0) atomicAdd(dst + 192, static_cast<unsigned long long>(fs));
                if (y == a.H - 1) atomicAdd(dst + 256, static_cast<unsigned long long>(fs));
                if (seg == 0) {
                  const long long f0 = __float2ll_rn(pool_g[256 + et] * kFx);
                  fx_c0 += f0;
                  if (y == 0) atomicAdd(dst + 320, static_cast<unsigned long long>(f0));          // K00
                  if (y == a.H - 1) atomicAdd(dst + 448, static_cast<unsigned long long>(f0));    // KH0
                }
                if (seg == nseg - 1) {
                  const long long fl = __float2ll_rn(pool_g[320 + et] * kFx);
                  fx_cl += fl;
                  if (y == 0) atomicAdd(dst + 384, static_cast<unsigned long long>(fl));          // K0W
                  if (y == a.H - 1) atomicAdd(dst + 512, static_cast<unsigned long long>(fl));    // KHW
                }
              } else {
                a.pool_rows[(static_cast<size_t>(col) * a.H + y) * 64 + et] = s;
              }
            }
          }
          if (q == 0) DFIR_TRACE(14 + egrp, it >> (kTwoEpi ? 1 : 0));
        }
      }
      if constexpr (EPI == EPI_RELU_STATS) {
        if (a.istats != nullptr && et < 64 && cur_img >= 0) {
          unsigned long long* dst = reinterpret_cast<unsigned long long*>(a.istats) + static_cast<size_t>(cur_img) * 576 + et;
          atomicAdd(dst, static_cast<unsigned long long>(fx_t));
          atomicAdd(dst + 64, static_cast<unsigned long long>(fx_c0));
          atomicAdd(dst + 128, static_cast<unsigned long long>(fx_cl));
        }
      }
      if (EPI != EPI_TAIL_NCHW && EPI != EPI_SCALE_SKIP && EPI != EPI_RELU_STATS_W && !kHL && et == 0) tma_store_wait<0>();
      if ((kHL || EPI == EPI_RELU_STATS_W) && lane == 0) tma_store_wait<0>();
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

int make_tmap_nhwc_bf16(CUtensorMap* m, const void* base, int C, int W, int H, int B, long long pix_stride_bytes,
                        long long row_stride_bytes, long long img_stride_bytes, int box_w) {
  PFN_encodeTiled enc = get_encode();
  if (enc == nullptr) return DFIR_ERR_DRIVER;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(B)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(pix_stride_bytes), static_cast<cuuint64_t>(row_stride_bytes),
                           static_cast<cuuint64_t>(img_stride_bytes)};
  cuuint32_t box[4] = {64u, static_cast<cuuint32_t>(box_w), 1u, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? DFIR_OK : DFIR_ERR_TMAP;
}

// dense NHWC uint8 plane (the 8-bit lo plane of the residual stream): box = 64 channels x box_w pixels, SWIZZLE_64B
int make_tmap_nhwc_u8(CUtensorMap* m, const void* base, int C, int W, int H, int B, int box_w) {
  PFN_encodeTiled enc = get_encode();
  if (enc == nullptr) return DFIR_ERR_DRIVER;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(B)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(C) * W, static_cast<cuuint64_t>(C) * W * H};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(C), static_cast<cuuint32_t>(box_w), 1u, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? DFIR_OK : DFIR_ERR_TMAP;
}

template <int NT, int EPI, int INMODE>
static int launch_one(const CUtensorMap& tin, const CUtensorMap& tout, const ConvHlMaps& hl, const ConvTcArgs& a, int grid,
                      cudaStream_t stream) {
  using L = SmemLayout<NT>;
  static bool configured[64] = {};  // per device: the attribute lives in the device's context
  auto kern = conv3x3_c64_tc_kernel<NT, EPI, INMODE>;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return DFIR_ERR_CUDA;
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total + 1024) != cudaSuccess)
      return DFIR_ERR_CUDA;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  static const bool use_pdl = getenv("DFIR_PDL") == nullptr || (atoi(getenv("DFIR_PDL")) & 1) != 0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(conv_threads<EPI, INMODE>());
  cfg.dynamicSmemBytes = L::total + 1024;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, tin, tout, hl, a) == cudaSuccess ? DFIR_OK : DFIR_ERR_CUDA;
}

int debug_trace(unsigned long long* out1024) {
  if (cudaDeviceSynchronize() != cudaSuccess) return DFIR_ERR_CUDA;
  return cudaMemcpyFromSymbol(out1024, g_dfir_trace, sizeof(unsigned long long) * 16 * 64) == cudaSuccess ? DFIR_OK
                                                                                                           : DFIR_ERR_CUDA;
}

int debug_watchdog(unsigned int* out8, int reset) {
  if (cudaMemcpyFromSymbol(out8, g_dfir_watchdog, 8 * sizeof(unsigned int)) != cudaSuccess) return DFIR_ERR_CUDA;
  if (out8[0] == 0u) {  // nothing recorded by the conv: report the weight-gradient kernel's record (tags 21-26)
    if (wgrad_tc_watchdog(out8, reset) != DFIR_OK) return DFIR_ERR_CUDA;
  } else if (reset) {
    unsigned int dummy[8];
    wgrad_tc_watchdog(dummy, reset);
  }
  if (reset == 2) {  // debugging: also print the block-0 progress probes
    unsigned int pg[16];
    if (cudaMemcpyFromSymbol(pg, g_dfir_progress, sizeof(pg)) == cudaSuccess) {
      printf("progress:");
      for (int i = 0; i < 16; ++i) printf(" %u", pg[i]);
      printf("\n");
    }
  }
  if (reset) {
    unsigned int z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (cudaMemcpyToSymbol(g_dfir_watchdog, z, sizeof(z)) != cudaSuccess) return DFIR_ERR_CUDA;
  }
  return DFIR_OK;
}

int conv3x3_c64_tc(const ConvTcDesc& d, cudaStream_t stream) {
  if (d.B <= 0 || d.H <= 0 || d.W <= 0) return DFIR_OK;
  if (d.cin_total % 64 != 0 || d.cin_off % 64 != 0) return DFIR_ERR_ARG;
  const bool fused = d.in_mode == IN_FUSED;
  if (fused && (d.r_bf16 == nullptr || d.xin_f32 == nullptr || d.cin_total != 64 || d.xin_f32 == d.xout_f32))
    return DFIR_ERR_ARG;
  if (fused && d.ca_style != DFIR_STYLE_NONE &&
      (d.pool_rows == nullptr || d.ca_params == nullptr || d.ca_A > 512 || d.ca_M > 448 ||
       (d.ca_A > 0 && d.attributes == nullptr)))
    return DFIR_ERR_ARG;
  if (!fused && d.in_bf16 == nullptr) return DFIR_ERR_ARG;
  if (d.epi == EPI_SCALE_SKIP && (d.out_bf16 == nullptr || d.out_pix_stride != 128 ||
                                  d.out_row_stride != static_cast<long long>(d.W) * 128))
    return DFIR_ERR_ARG;  // the direct-store epilogue writes dense NHWC
  const bool hl_mode = d.epi == EPI_SCALE_SKIP_HL || d.epi == EPI_SCALE_SKIP_HL8;
  const bool lo8 = d.epi == EPI_SCALE_SKIP_HL8;
  if (hl_mode && (fused || d.out_bf16 == nullptr || d.skip_hi == nullptr || d.skip_lo == nullptr || d.out_pix_stride != 128 ||
                  d.out_row_stride != static_cast<long long>(d.W) * 128 || d.r_out != nullptr || (d.relu_out && !lo8)))
    return DFIR_ERR_ARG;  // the stream planes are dense NHWC; training extras live on the fp32-stream epilogue
  if (d.epi_stats == 2 && (!hl_mode || d.istats == nullptr)) return DFIR_ERR_ARG;
  if ((d.epi == EPI_SCALE_SKIP || hl_mode) && d.epi_stats &&
      (fused || d.ca_style == DFIR_STYLE_NONE || (d.epi_stats != 2 && (d.pool_rows == nullptr || d.col_first == nullptr || d.col_last == nullptr)) || d.ca_params == nullptr ||
       d.ca_A > 512 || d.ca_M > 448 || (d.ca_A > 0 && d.attributes == nullptr)))
    return DFIR_ERR_ARG;
  if (d.epi == EPI_RELU_MASK && d.mask_bf16 == nullptr) return DFIR_ERR_ARG;
  if (d.epi == EPI_RELU_STATS && d.istats == nullptr && (d.pool_rows == nullptr || d.col_first == nullptr || d.col_last == nullptr))
    return DFIR_ERR_ARG;
  // (a thread of the warp-autonomous statistics epilogue keeps ONE column accumulator: it must not own both the first and
  // the last column, i.e. the last pixel of the last row segment must not fall on pixel 0, 8, 16 or 24 of warp 0)
  if (d.epi == EPI_RELU_STATS_W && (d.istats == nullptr || fused || ((d.W - 1) % 128 < 32 && ((d.W - 1) % 128) % 8 == 0)))
    return DFIR_ERR_ARG;
  CUtensorMap tin, tout;
  int rc = DFIR_OK;
  if (!fused) {
    // dense NHWC unless explicit byte strides are given (backward of the upsampler: one sub-pixel phase of a
    // PixelShuffle output is a strided view, advanced/common.py:30)
    const long long ips = d.in_pix_stride > 0 ? d.in_pix_stride : static_cast<long long>(d.cin_total) * 2;
    const long long irs = d.in_row_stride > 0 ? d.in_row_stride : static_cast<long long>(d.W) * d.cin_total * 2;
    const long long iis = d.in_img_stride > 0 ? d.in_img_stride : static_cast<long long>(d.H) * d.W * d.cin_total * 2;
    rc = make_tmap_nhwc_bf16(&tin, d.in_bf16, d.cin_total, d.W, d.H, d.B, ips, irs, iis, kBoxPix);
    if (rc != DFIR_OK) return rc;
  }
  if (d.epi != EPI_TAIL_NCHW) {
    rc = make_tmap_nhwc_bf16(&tout, d.out_bf16, 64, d.W, d.H, d.B, d.out_pix_stride, d.out_row_stride,
                             d.out_img_stride, 128);
    if (rc != DFIR_OK) return rc;
    if (fused) tin = tout;  // unused by the kernel in IN_FUSED mode, but must be a valid descriptor
  } else {
    if (fused) return DFIR_ERR_ARG;
    tout = tin;
  }
  ConvHlMaps hl{};
  if (hl_mode) {
    const long long rowB = static_cast<long long>(d.W) * 128, imgB = rowB * d.H;
    const void* planes[4] = {d.skip_hi, d.skip_lo, d.out_bf16, d.out_lo != nullptr ? d.out_lo : d.out_bf16};
    for (int i = 0; i < 4; ++i) {
      const bool lo_plane = (i & 1) != 0 && !(i == 3 && d.out_lo == nullptr);
      rc = (lo8 && lo_plane) ? make_tmap_nhwc_u8(&hl.m[i], planes[i], 64, d.W, d.H, d.B, 32)
                             : make_tmap_nhwc_bf16(&hl.m[i], planes[i], 64, d.W, d.H, d.B, 128, rowB, imgB, lo8 ? 32 : 16);
      if (rc != DFIR_OK) return rc;
    }
  } else {
    for (int i = 0; i < 4; ++i) hl.m[i] = tout;  // unused, but must be valid descriptors
    if (d.epi == EPI_RELU_STATS_W) {  // the warps store 32 pixels each
      rc = make_tmap_nhwc_bf16(&hl.m[0], d.out_bf16, 64, d.W, d.H, d.B, d.out_pix_stride, d.out_row_stride, d.out_img_stride, 32);
      if (rc != DFIR_OK) return rc;
    }
  }
  ConvTcArgs a{};
  a.hl_store_lo = d.out_lo != nullptr ? 1 : 0;
  a.flip = (hl_mode && d.flip && d.W <= 128) ? 1 : 0;
  a.istats = d.istats;
  a.istats_clear = d.istats_clear;
  a.B = d.B;
  a.H = d.H;
  a.W = d.W;
  a.nseg = (d.W + 127) / 128;
  a.cin_off = d.cin_off;
  a.cout = d.cout;
  a.wpacked = d.wpacked;
  a.bias = d.bias;
  a.skip_f32 = d.skip_f32;
  a.out_f32 = d.out_f32;
  a.pool_rows = d.pool_rows;
  a.col_first = d.col_first;
  a.col_last = d.col_last;
  a.svec = d.svec;
  a.mask_bf16 = reinterpret_cast<const __nv_bfloat16*>(d.mask_bf16);
  a.r_out = reinterpret_cast<__nv_bfloat16*>(d.r_out);
  a.ymean_out = d.ymean_out;
  a.relu_out = d.relu_out;
  a.tail_accumulate = d.tail_accumulate;
  a.out_bf16_direct = reinterpret_cast<__nv_bfloat16*>(d.out_bf16);
  a.r_bf16 = reinterpret_cast<const __nv_bfloat16*>(d.r_bf16);
  a.xin_f32 = d.xin_f32;
  a.xout_f32 = d.xout_f32;
  a.res_scale = d.res_scale;
  a.epi_stats = d.epi_stats;
  a.debug_probe = getenv("DFIR_DEBUG_PROBE") != nullptr ? atoi(getenv("DFIR_DEBUG_PROBE")) : 0;
  {
    const char* e = getenv("DFIR_WPREFETCH");  // next-layer weight prefetch: measured, no gain (31.6-32.7 ms either way), default off
    if (d.next_wpacked != nullptr && e != nullptr && atoi(e) != 0) {
      a.next_w = d.next_wpacked;
      a.next_w_bytes = 9 * 64 * 128;
    }
  }
  {
    // L2 residency of the RCAB chain: t (conv1 -> conv2) and the bf16 copy of the stream (conv2 -> conv1) are consumed
    // by the next launch and fit the 126 MB L2 together, the fp32 stream (read and written once per block) does not.
    // DFIR_L2_POLICY = six letters n|f|l (normal / evict first / evict last):
    //   conv1 input, conv1 output, conv2 input, conv2 bf16 output, conv2 fp32 skip, conv2 fp32 output
    const char* pol_env = getenv("DFIR_L2_POLICY");
    const char* pol = pol_env != nullptr ? pol_env : kDefaultL2Policy;
    auto word = [](char c) -> unsigned long long {
      return c == 'f' ? ptx::kL2EvictFirst : (c == 'l' ? ptx::kL2EvictLast : ptx::kL2EvictNormal);
    };
    const bool conv1_role = d.epi == EPI_RELU_STATS || d.epi == EPI_RELU_STATS_W || d.epi == EPI_BIAS_RELU;
    const bool conv2_role = d.epi == EPI_SCALE_SKIP || hl_mode;  // HL: letters 5, 6 = lo plane in / out
    if (pol != nullptr && strlen(pol) >= 6 && strncmp(pol, "nnnnnn", 6) != 0 && !fused && (conv1_role || conv2_role)) {
      a.use_hints = 1;
      a.pol_in = word(pol[conv1_role ? 0 : 2]);
      a.pol_out = word(pol[conv1_role ? 1 : 3]);
      a.pol_skip = word(pol[4]);
      a.pol_f32 = word(pol[5]);
    }
  }
  a.ca_params = d.ca_params;
  a.attributes = d.attributes;
  a.sq = d.sq;
  a.ca_style = d.ca_style;
  a.ca_R = d.ca_R;
  a.ca_M = d.ca_M;
  a.ca_A = d.ca_A;
  const long long G = static_cast<long long>(d.B) * a.nseg * d.H;
  int grid = d.num_sms > 0 ? d.num_sms : 148;
  if (const char* e = getenv("DFIR_NUM_SMS")) grid = atoi(e) > 0 ? atoi(e) : grid;  // debugging aid
  if (G < grid) grid = static_cast<int>(G);
  // IN_FUSED keeps the attention vectors of every image a band touches in shared memory
  if ((fused || ((d.epi == EPI_SCALE_SKIP || hl_mode) && d.epi_stats)) &&
      (G + grid - 1) / grid > static_cast<long long>(kMaxBandImages - 1) * a.nseg * d.H)
    return DFIR_ERR_ARG;
  if (fused) {
    switch (d.epi) {
      case EPI_BIAS_RELU: return launch_one<64, EPI_BIAS_RELU, IN_FUSED>(tin, tout, hl, a, grid, stream);
      case EPI_BIAS_SKIP: return launch_one<64, EPI_BIAS_SKIP, IN_FUSED>(tin, tout, hl, a, grid, stream);
      case EPI_SCALE_SKIP: return launch_one<64, EPI_SCALE_SKIP, IN_FUSED>(tin, tout, hl, a, grid, stream);
      default: return DFIR_ERR_ARG;
    }
  }
  switch (d.epi) {
    case EPI_BIAS: return launch_one<64, EPI_BIAS, IN_TMA>(tin, tout, hl, a, grid, stream);
    case EPI_BIAS_RELU: return launch_one<64, EPI_BIAS_RELU, IN_TMA>(tin, tout, hl, a, grid, stream);
    case EPI_BIAS_POOL: return launch_one<64, EPI_BIAS_POOL, IN_TMA>(tin, tout, hl, a, grid, stream);
    case EPI_BIAS_SKIP: return launch_one<64, EPI_BIAS_SKIP, IN_TMA>(tin, tout, hl, a, grid, stream);
    case EPI_RELU_STATS: return launch_one<64, EPI_RELU_STATS, IN_TMA>(tin, tout, hl, a, grid, stream);
    case EPI_RELU_STATS_W: return launch_one<64, EPI_RELU_STATS_W, IN_TMA>(tin, tout, hl, a, grid, stream);
    case EPI_RELU_MASK: return launch_one<64, EPI_RELU_MASK, IN_TMA>(tin, tout, hl, a, grid, stream);
    case EPI_SCALE_SKIP: return launch_one<64, EPI_SCALE_SKIP, IN_TMA>(tin, tout, hl, a, grid, stream);
    case EPI_TAIL_NCHW: return launch_one<16, EPI_TAIL_NCHW, IN_TMA>(tin, tout, hl, a, grid, stream);
    case EPI_SCALE_SKIP_HL: return launch_one<64, EPI_SCALE_SKIP_HL, IN_TMA>(tin, tout, hl, a, grid, stream);
    case EPI_SCALE_SKIP_HL8: return launch_one<64, EPI_SCALE_SKIP_HL8, IN_TMA>(tin, tout, hl, a, grid, stream);
    default: return DFIR_ERR_ARG;
  }
}

}  // namespace dfir
