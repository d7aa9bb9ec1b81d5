// 3x3 convolution, 64 input channels, as an implicit GEMM on the sm_100a tensor cores.
//
//   reference op : default_conv = nn.Conv2d(k=3, stride 1, zero pad 1, bias)
//                  (/root/reference/Code/SISR/models/advanced/common.py:5-8)
//
// Data layout (HBM): activations NHWC bf16 (one pixel = 64 ch = 128 B = exactly one 128B-swizzle row),
// weights pre-packed per tap as [tap][cout][cin] bf16 in the UMMA K-major SWIZZLE_128B byte order.
//
// GEMM view: D[M = 128 pixels of one image row][N = cout] += A[M][K = 64 cin] * B[N][K] for each of the 9
// taps (K total = 576).  No im2col is ever materialised: a persistent CTA streams input rows into a shared
// memory ring with TMA (one box = one row segment of 130 pixels incl. the x halo; out-of-bounds pixels and
// rows are zero-filled by the TMA unit, which IS the conv's zero padding) and every tap's A operand is
// just a shifted view of a ring slot: dy picks the slot, dx shifts the descriptor start by dx*128 B
// (base_offset stays 0: the hardware derives the swizzle phase from the absolute smem address).  The weights of the layer (9 x cout x 128 B) stay resident in
// shared memory for the CTA's whole life.  Accumulators live in TMEM (4 buffers) so the epilogue of row i
// overlaps the MMAs of rows i+1..i+3.
//
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + tcgen05.mma issuer, warps 2-5 = epilogue
// (TMEM -> registers -> fused bias/ReLU/skip/pool -> swizzled smem -> TMA store).
//
// Input modes
//   IN_TMA   : the input rows are bf16 NHWC in HBM/L2 and arrive by TMA (192 threads).
//   IN_FUSED : the input of the conv is x' = r * s + x, i.e. the previous block's channel-attention scale
//              and residual add (QRCAB: `res * y` twice and `res += x`, architectures.py:127,172-180;
//              q_layer.py:43).  Eight extra warps (two groups that alternate rows) read r (bf16) and x
//              (fp32 residual stream), form x', write it back as the new fp32 stream for the rows the CTA
//              owns, and deposit bf16(x') straight into the swizzled ring slot the tensor core reads.  The
//              separate elementwise pass (and its 12 B/element of traffic) disappears (448 threads).
// Attention tail (EPI_BIAS_POOL with a.svec_out): the last CTA to finish an image's rows turns the per-row
// pooled sums into the block's attention vector s = CA(mean) * meta_scale (attn.cuh) for that image.
#include "ptx.cuh"
#include "kernels.h"
#include "attn.cuh"

#include <cuda_bf16.h>

namespace dfir {

using namespace ptx;

namespace {

constexpr int kSlots = 6;                    // input-row ring depth
constexpr int kSlotPix = 136;                // 130 px used (128 + 2 halo), rounded up to 8 px = 1024 B
constexpr int kSlotBytes = kSlotPix * 128;   // 17408, multiple of 1024
constexpr int kBoxPix = 130;
constexpr int kRowBytes = kBoxPix * 128;     // bytes one TMA row load delivers
constexpr int kAcc = 4;                      // TMEM accumulator buffers
constexpr int kStageBytes = 128 * 128;       // one output row segment, bf16
constexpr int kThreads = 192;            // IN_TMA
constexpr int kThreadsFused = 192 + 256;  // + two transform groups of 4 warps

template <int NT>
struct SmemLayout {
  static constexpr int w_bytes = 9 * NT * 128;
  static constexpr int off_w = 0;
  static constexpr int off_ring = off_w + ((w_bytes + 1023) / 1024) * 1024;
  static constexpr int off_stage = off_ring + kSlots * kSlotBytes;
  static constexpr int off_bias = off_stage + 2 * kStageBytes;
  static constexpr int off_pool = off_bias + 64 * 4;
  static constexpr int off_attn = off_pool + 4 * 64 * 4;          // y[64] s[64] attr[512] tmp[1024] flag
  static constexpr int off_svec = off_attn + (64 + 64 + 512 + 1024 + 4) * 4;  // s of the image(s) in flight, 2x64
  static constexpr int off_bars = off_svec + 2 * 64 * 4;
  static constexpr int n_bars = 2 * kSlots + 2 * kAcc + 1;
  static constexpr int off_tmem = off_bars + n_bars * 8;
  static constexpr int total = off_tmem + 16;
};

// butterfly transpose-reduce: on entry lane l holds v[0..63] (channel values of its pixel, already masked);
// on exit v[0], v[1] hold the sums over the warp's 32 pixels of channels 2*l and 2*l+1.
__device__ __forceinline__ void warp_channel_sums(float (&v)[64], int lane) {
#pragma unroll
  for (int step = 0; step < 5; ++step) {
    const int n = 64 >> step;       // values held before this step
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

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

}  // namespace

template <int NT, int EPI, int INMODE>
__global__ void __launch_bounds__(INMODE == IN_FUSED ? kThreadsFused : kThreads, 1)
conv3x3_c64_tc_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_out,
                      ConvTcArgs a) {
  using L = SmemLayout<NT>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms (TMA and UMMA) need 1024-byte alignment of every tile base
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* wsm = smem + L::off_w;
  uint8_t* ring = smem + L::off_ring;
  uint8_t* stage = smem + L::off_stage;
  float* bias_s = reinterpret_cast<float*>(smem + L::off_bias);
  float* pool_s = reinterpret_cast<float*>(smem + L::off_pool);
  float* attn_s = reinterpret_cast<float*>(smem + L::off_attn);
  float* svec_s = reinterpret_cast<float*>(smem + L::off_svec);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::off_bars);
  uint64_t* full = bars;
  uint64_t* empty = bars + kSlots;
  uint64_t* tfull = bars + 2 * kSlots;
  uint64_t* tempty = bars + 2 * kSlots + kAcc;
  uint64_t* wbar = bars + 2 * kSlots + 2 * kAcc;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + L::off_tmem);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

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

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_in);
    if (EPI != EPI_TAIL_NCHW) prefetch_tmap(&tmap_out);
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(&full[i], INMODE == IN_FUSED ? 128 : 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < kAcc; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    mbar_init(wbar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + NT) bias_s[threadIdx.x - 64] = a.bias[threadIdx.x - 64];
  if (warp == 1) tmem_alloc<kAcc * NT>(tmem_holder);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (g0 < g1) {
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (elect_one()) {
        mbar_arrive_expect_tx(wbar, L::w_bytes);
        bulk_load_1d(wsm, a.wpacked, L::w_bytes, wbar);
        const int pr_last = INMODE == IN_FUSED ? pr_first - 1 : padded(g1 - 1) + 1;
        for (int pr = pr_first, n = 0; pr <= pr_last; ++pr, ++n) {
          const int slot = n % kSlots;
          const uint32_t use = n / kSlots;
          mbar_wait(&empty[slot], (use & 1) ^ 1);
          const int col = pr / Hp;
          const int yy = pr % Hp - 1;
          const int b = col / nseg;
          const int seg = col % nseg;
          mbar_arrive_expect_tx(&full[slot], kRowBytes);
          tma_load_4d(ring + slot * kSlotBytes, &tmap_in, &full[slot], a.cin_off, seg * 128 - 1, yy, b);
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      // The whole warp runs the loop so that control flow and descriptor arithmetic stay warp-uniform
      // (uniform registers feed UTCHMMA directly); only the elected lane issues tcgen05 instructions.
      const bool leader = elect_one();
      constexpr uint32_t idesc = make_idesc_bf16_f32(128, NT);
      const uint64_t da_base = make_sw128_kmajor_desc(smem_u32(ring), 1024, 0);
      const uint64_t db_base = make_sw128_kmajor_desc(smem_u32(wsm), 1024, 0);
      mbar_wait(wbar, 0);
      int released = 0;  // next ring sequence index to hand back to the producer
      for (int g = g0, it = 0; g < g1; ++g, ++it) {
        const int nc = padded(g) - pr_first;  // ring sequence index of the centre row
        const int acc = it % kAcc;
        mbar_wait(&tempty[acc], (((it / kAcc) & 1) ^ 1));
        uint32_t slot_off[3];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const int n = nc - 1 + dy;
          const int slot = n % kSlots;
          mbar_wait(&full[slot], (n / kSlots) & 1);
          slot_off[dy] = static_cast<uint32_t>(slot * (kSlotBytes >> 4));
        }
        tcgen05_fence_after();
        if (leader) {
          const uint32_t d_tmem = tmem_base + acc * NT;
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const uint64_t da_row = da_base + slot_off[dy];
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                // start-address field is in 16-byte units: +8 per pixel (dx), +2 per 16-channel k-step
                const uint64_t da = da_row + static_cast<uint32_t>(dx * 8 + k * 2);
                const uint64_t db = db_base + static_cast<uint32_t>(((dy * 3 + dx) * NT * 128 + k * 32) >> 4);
                umma_f16_ss(d_tmem, da, db, idesc, (dy | dx | k) != 0 ? 1u : 0u);
              }
            }
          }
          umma_commit(&tfull[acc]);
          for (int rel = released; rel <= nc - 1; ++rel) umma_commit(&empty[rel % kSlots]);
        }
        released = nc > released ? nc : released;
        __syncwarp();
      }
    } else if (warp >= 6) {
      // ===================== fused input transform (IN_FUSED only; warps 6..13) =====================
      if constexpr (INMODE == IN_FUSED) {
        const int tt = threadIdx.x - kThreads;  // 0..255
        const int tg = tt >> 7;                 // group: handles ring sequence indices n = tg (mod 2)
        const int tl = tt & 127;
        float* s_loc = svec_s + tg * 64;
        const int n_last = padded(g1 - 1) + 1 - pr_first;
        const int pc_first = padded(g0), pc_last = padded(g1 - 1);
        int cur_b = -1;
        for (int n = tg; n <= n_last; n += 2) {
          const int pr = pr_first + n;
          const int slot = n % kSlots;
          const int col = pr / Hp;
          const int yy = pr % Hp - 1;
          const int b = col / nseg;
          const int seg = col % nseg;
          const bool row_ok = yy >= 0 && yy < H;
          const bool owned = row_ok && pr >= pc_first && pr <= pc_last;
          if (row_ok && b != cur_b) {  // (group-uniform) fetch this image's attention vector
            named_bar_sync(3 + tg, 128);
            if (tl < 64) s_loc[tl] = a.svec_in[static_cast<size_t>(b) * 64 + tl];
            named_bar_sync(3 + tg, 128);
            cur_b = b;
          }
          mbar_wait(&empty[slot], ((n / kSlots) & 1) ^ 1);
          uint8_t* srow = ring + slot * kSlotBytes;
          for (int p = tl; p < kBoxPix; p += 128) {
            const int x = seg * 128 - 1 + p;
            uint4* dst = reinterpret_cast<uint4*>(srow + p * 128);
            if (row_ok && x >= 0 && x < a.W) {
              const size_t e = ((static_cast<size_t>(b) * H + yy) * a.W + x) * 64;
              const bool wr = owned && p >= 1 && p <= 128 && a.xout_f32 != nullptr;
#pragma unroll
              for (int h = 0; h < 2; ++h) {  // two halves of 32 channels bound the register footprint
                uint4 rr[4];
                float4 xx[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) rr[i] = *reinterpret_cast<const uint4*>(a.r_bf16 + e + h * 32 + i * 8);
#pragma unroll
                for (int i = 0; i < 8; ++i) xx[i] = *reinterpret_cast<const float4*>(a.xin_f32 + e + h * 32 + i * 4);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&rr[i]);
                  const float* sc = s_loc + h * 32 + i * 8;
                  float o[8];
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float2 f = __bfloat1622float2(hb[j]);
                    const float4 xv = xx[2 * i + (j >> 1)];
                    const float xa = (j & 1) ? xv.z : xv.x;
                    const float xb = (j & 1) ? xv.w : xv.y;
                    o[2 * j] = fmaf(f.x, sc[2 * j], xa);
                    o[2 * j + 1] = fmaf(f.y, sc[2 * j + 1], xb);
                  }
                  if (wr) {
                    float4* xo = reinterpret_cast<float4*>(a.xout_f32 + e + h * 32 + i * 8);
                    xo[0] = make_float4(o[0], o[1], o[2], o[3]);
                    xo[1] = make_float4(o[4], o[5], o[6], o[7]);
                  }
                  uint4 pk;
                  pk.x = pack_bf16x2(o[0], o[1]);
                  pk.y = pack_bf16x2(o[2], o[3]);
                  pk.z = pack_bf16x2(o[4], o[5]);
                  pk.w = pack_bf16x2(o[6], o[7]);
                  dst[(h * 4 + i) ^ (p & 7)] = pk;
                }
              }
            } else {
              const uint4 z = make_uint4(0u, 0u, 0u, 0u);  // zero padding (rows -1 / H, columns -1 / W)
#pragma unroll
              for (int c = 0; c < 8; ++c) dst[c] = z;
            }
          }
          fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core's async proxy
          mbar_arrive(&full[slot]);
        }
      }
    } else {
      // ===================== epilogue (warps 2..5) =====================
      const int q = warp & 3;          // TMEM lane quarter this warp may read
      const int m = q * 32 + lane;     // pixel within the 128-px row segment
      const int et = threadIdx.x - 64; // 0..127
      for (int g = g0, it = 0; g < g1; ++g, ++it) {
        const int col = g / H;
        const int y = g % H;
        const int b = col / nseg;
        const int seg = col % nseg;
        const int x = seg * 128 + m;
        const bool valid = x < a.W;
        const int acc = it % kAcc;
        mbar_wait(&tfull[acc], (it / kAcc) & 1);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * NT;
        float v[NT];
        if constexpr (NT == 64) {
          uint32_t r0[32], r1[32];
          tmem_ld_32x32b_x32(taddr, r0);
          tmem_ld_32x32b_x32(taddr + 32, r1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v[i] = __uint_as_float(r0[i]);
            v[32 + i] = __uint_as_float(r1[i]);
          }
        } else {
          uint32_t r0[16];
          tmem_ld_32x32b_x16(taddr, r0);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r0[i]);
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);

#pragma unroll
        for (int i = 0; i < NT; ++i) v[i] += bias_s[i];

        if constexpr (EPI == EPI_TAIL_NCHW) {
          // fp32 NCHW output, a.cout real channels (<= NT)
          if (valid) {
            const size_t plane = static_cast<size_t>(a.H) * a.W;
            float* o = a.out_f32 + (static_cast<size_t>(b) * a.cout) * plane + static_cast<size_t>(y) * a.W + x;
#pragma unroll
            for (int c = 0; c < NT; ++c)
              if (c < a.cout) o[c * plane] = v[c];
          }
        } else {
          if constexpr (EPI == EPI_BIAS_RELU) {
#pragma unroll
            for (int i = 0; i < NT; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          if constexpr (EPI == EPI_BIAS_SKIP) {
            if (valid) {
              const size_t pix = (static_cast<size_t>(b) * a.H + y) * a.W + x;
              const float4* sk = reinterpret_cast<const float4*>(a.skip_f32 + pix * 64);
              float4* o = reinterpret_cast<float4*>(a.out_f32 + pix * 64);
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                const float4 s = sk[c];
                v[4 * c + 0] += s.x;
                v[4 * c + 1] += s.y;
                v[4 * c + 2] += s.z;
                v[4 * c + 3] += s.w;
                if (a.out_f32 != nullptr) o[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
              }
            }
          }
          // ---- bf16 row segment -> swizzled staging -> TMA store
          const int sb = it & 1;
          uint8_t* st = stage + sb * kStageBytes;
          if (et == 0) tma_store_wait_read<1>();
          named_bar_sync(1, 128);
          {
            uint4* row = reinterpret_cast<uint4*>(st + m * 128);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              uint4 pk;
              pk.x = pack_bf16x2(v[8 * c + 0], v[8 * c + 1]);
              pk.y = pack_bf16x2(v[8 * c + 2], v[8 * c + 3]);
              pk.z = pack_bf16x2(v[8 * c + 4], v[8 * c + 5]);
              pk.w = pack_bf16x2(v[8 * c + 6], v[8 * c + 7]);
              row[c ^ (m & 7)] = pk;
            }
          }
          if constexpr (EPI == EPI_BIAS_POOL) {
            // per-row channel sums of the fp32 (pre-rounding) conv output, garbage pixels masked
            if (!valid) {
#pragma unroll
              for (int i = 0; i < NT; ++i) v[i] = 0.f;
            }
            warp_channel_sums(v, lane);
            pool_s[q * 64 + 2 * lane] = v[0];
            pool_s[q * 64 + 2 * lane + 1] = v[1];
          }
          fence_proxy_async_smem();
          named_bar_sync(2, 128);
          if (et == 0) {
            tma_store_4d(&tmap_out, st, 0, seg * 128, y, b);
            tma_store_commit();
          }
          if constexpr (EPI == EPI_BIAS_POOL) {
            if (et < 64) {
              const float s = ((pool_s[et] + pool_s[64 + et]) + pool_s[128 + et]) + pool_s[192 + et];
              a.pool_rows[(static_cast<size_t>(col) * a.H + y) * 64 + et] = s;
            }
            // ---- attention tail: when this CTA has finished its share of image b, count it in; the last
            // CTA of the image reduces the pooled rows (fixed order) and evaluates the attention vector.
            const bool img_done = (g + 1 == g1) || (((g + 1) / H) / nseg != b);
            if (a.svec_out != nullptr && img_done) {
              const int rows_img = nseg * H;
              int* flag = reinterpret_cast<int*>(attn_s + 64 + 64 + 512 + 1024);
              __threadfence();
              named_bar_sync(1, 128);
              if (et == 0) {
                const long long ga = static_cast<long long>(b) * rows_img, gb = ga + rows_img - 1;
                const int i_first = static_cast<int>(((ga + 1) * gridDim.x - 1) / G);
                const int i_last = static_cast<int>(((gb + 1) * gridDim.x - 1) / G);
                const int old = atomicAdd(a.img_counter + b, 1);
                const int last = old == (i_last - i_first);
                if (last) a.img_counter[b] = 0;  // self-resetting: ready for the next layer
                *flag = last;
              }
              named_bar_sync(2, 128);
              if (*flag) {
                __threadfence();
                float* y_s = attn_s;
                float* s_s = attn_s + 64;
                float* attr_s = attn_s + 128;
                float* tmp = attn_s + 128 + 512;
                const NamedGroup grp{et, 128, 5};
                {
                  const int c = et & 63, half = et >> 6;
                  const float* pr = a.pool_rows + static_cast<size_t>(b) * rows_img * 64 + c;
                  float sum = 0.f;
                  for (int row = half; row < rows_img; row += 2) sum += pr[static_cast<size_t>(row) * 64];
                  tmp[half * 64 + c] = sum;
                }
                for (int i = et; i < a.ca_A; i += 128) attr_s[i] = a.attributes[static_cast<size_t>(b) * a.ca_A + i];
                grp.sync();
                if (et < 64) y_s[et] = (tmp[et] + tmp[64 + et]) / (static_cast<float>(H) * static_cast<float>(a.W));
                grp.sync();
                attn_vector(grp, a.ca_style, a.ca_params, 64, a.ca_R, a.ca_M, attr_s, y_s, s_s, tmp);
                if (et < 64)
                  a.svec_out[static_cast<size_t>(b) * 64 + et] =
                      s_s[et] * (a.sq != nullptr ? a.sq[static_cast<size_t>(b) * 64 + et] : 1.f);
              }
              named_bar_sync(1, 128);  // keep *flag / attn scratch stable until everyone has read it
            }
          }
        }
      }
      if (EPI != EPI_TAIL_NCHW && et == 0) tma_store_wait<0>();
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<kAcc * NT>(tmem_base);
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

template <int NT, int EPI, int INMODE>
static int launch_one(const CUtensorMap& tin, const CUtensorMap& tout, const ConvTcArgs& a, int grid,
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
  kern<<<grid, INMODE == IN_FUSED ? kThreadsFused : kThreads, L::total + 1024, stream>>>(tin, tout, a);
  return cudaGetLastError() == cudaSuccess ? DFIR_OK : DFIR_ERR_CUDA;
}

int conv3x3_c64_tc(const ConvTcDesc& d, cudaStream_t stream) {
  if (d.B <= 0 || d.H <= 0 || d.W <= 0) return DFIR_OK;
  if (d.cin_total % 64 != 0 || d.cin_off % 64 != 0) return DFIR_ERR_ARG;
  const bool fused = d.in_mode == IN_FUSED;
  if (fused && (d.r_bf16 == nullptr || d.xin_f32 == nullptr || d.svec_in == nullptr || d.cin_total != 64))
    return DFIR_ERR_ARG;
  if (!fused && d.in_bf16 == nullptr) return DFIR_ERR_ARG;
  CUtensorMap tin, tout;
  int rc = DFIR_OK;
  if (!fused) {
    rc = make_tmap_nhwc_bf16(&tin, d.in_bf16, d.cin_total, d.W, d.H, d.B, static_cast<long long>(d.cin_total) * 2,
                             static_cast<long long>(d.W) * d.cin_total * 2,
                             static_cast<long long>(d.H) * d.W * d.cin_total * 2, kBoxPix);
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
  ConvTcArgs a{};
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
  a.r_bf16 = reinterpret_cast<const __nv_bfloat16*>(d.r_bf16);
  a.xin_f32 = d.xin_f32;
  a.xout_f32 = d.xout_f32;
  a.svec_in = d.svec_in;
  a.svec_out = d.svec_out;
  a.img_counter = d.img_counter;
  a.ca_params = d.ca_params;
  a.attributes = d.attributes;
  a.sq = d.sq;
  a.ca_style = d.ca_style;
  a.ca_R = d.ca_R;
  a.ca_M = d.ca_M;
  a.ca_A = d.ca_A;
  if (d.svec_out != nullptr &&
      (d.epi != EPI_BIAS_POOL || d.img_counter == nullptr || d.ca_params == nullptr || d.ca_A > 512 || d.ca_M > 448))
    return DFIR_ERR_ARG;
  const long long G = static_cast<long long>(d.B) * a.nseg * d.H;
  int grid = d.num_sms > 0 ? d.num_sms : 148;
  if (G < grid) grid = static_cast<int>(G);
  if (fused) {
    switch (d.epi) {
      case EPI_BIAS_RELU: return launch_one<64, EPI_BIAS_RELU, IN_FUSED>(tin, tout, a, grid, stream);
      case EPI_BIAS_SKIP: return launch_one<64, EPI_BIAS_SKIP, IN_FUSED>(tin, tout, a, grid, stream);
      default: return DFIR_ERR_ARG;
    }
  }
  switch (d.epi) {
    case EPI_BIAS: return launch_one<64, EPI_BIAS, IN_TMA>(tin, tout, a, grid, stream);
    case EPI_BIAS_RELU: return launch_one<64, EPI_BIAS_RELU, IN_TMA>(tin, tout, a, grid, stream);
    case EPI_BIAS_POOL: return launch_one<64, EPI_BIAS_POOL, IN_TMA>(tin, tout, a, grid, stream);
    case EPI_BIAS_SKIP: return launch_one<64, EPI_BIAS_SKIP, IN_TMA>(tin, tout, a, grid, stream);
    case EPI_TAIL_NCHW: return launch_one<16, EPI_TAIL_NCHW, IN_TMA>(tin, tout, a, grid, stream);
    default: return DFIR_ERR_ARG;
  }
}

}  // namespace dfir
