// Weight gradient of the 64 -> 64 3x3 convolution on the sm_100a tensor cores (tcgen05, accumulators in TMEM):
//     dW[co][ci][dy][dx] = sum over pixels p of dY[p][co] * X[p + (dy-1, dx-1)][ci]      (zero padded)
//     db[co]             = sum over pixels p of dY[p][co]
// (backward of default_conv, /root/reference/Code/SISR/models/advanced/common.py:5-8, w.r.t. its parameters).
//
// GEMM view: the reduction (K) dimension is the PIXEL index, so both NHWC operands are MN-major: a shared-memory row
// is one pixel = 64 channels = 128 B, exactly what TMA delivers with the 128-byte swizzle.  MN-major SWIZZLE_128B
// canonical layout (in 16-byte units): ((8,n),(8,k)) : ((1,LBO),(8,SBO)) — 8 units = 64 channels along MN, K rows
// 128 B apart, SBO = 1024 B between groups of 8 pixels, LBO = distance between 64-channel atoms along MN.
//
// Trick: with LBO = 128 B the second MN atom of the A operand is the SAME X row shifted by one pixel, i.e. the next
// horizontal tap.  Per 16-pixel K step and per vertical tap dy the CTA issues
//     M=128 (taps dx=0,1 x 64 ci) x N=64 (co) x K=16      A = X row (y+dy-1) at pixel offset 0, LBO 128 B
//     M= 64 (tap  dx=2   x 64 ci) x N=64 (co) x K=16      A = the same row at pixel offset 2
// with B = the dY row.  Six accumulators (3 x 64 columns on 128 lanes + 3 x 64 columns on 64 lanes) hold the whole
// 9 x 64 x 64 gradient of the CTA's band of rows in tensor memory; they are written once at the end and reduced over
// CTAs by wgrad_reduce_kernel in a fixed order.
//
// Warp roles: 0 TMA producer (X rows with halo into a 5-slot ring, dY rows into 3 slots), 1 MMA issuer, 2-5 bias
// gradient while the rows stream (column sums of the dY tiles, 16-byte shared-memory loads), then the epilogue
// (TMEM -> partial sums in global memory).
#include "kernels.h"
#include "launch.cuh"
#include "ptx.cuh"

namespace dfir {

int make_tmap_nhwc_bf16(CUtensorMap* m, const void* base, int C, int W, int H, int B, long long pix_stride_bytes,
                        long long row_stride_bytes, long long img_stride_bytes, int box_w);

namespace {

using namespace ptx;

constexpr int kXS = 5;                    // X row ring
constexpr int kDS = 3;                    // dY row slots
constexpr int kXSlotB = 136 * 128;        // 130 px used, multiple of 1024
constexpr int kDSlotB = 128 * 128;
constexpr int kWtThreads = 192;            // producer, MMA, 4 x (bias gradient during the loop, then epilogue)
constexpr int kWtSmem = kXS * kXSlotB + kDS * kDSlotB + 256 + 16 * 64 * 4 + 1024;

struct WgradTcArgs {
  float* part;    // [grid][9][64 co][64 ci]
  float* dbpart;  // [grid][64]
  int B, H, W, nseg;
  int probe;  // DFIR_WGRAD_PROBE (timing experiments, wrong results): 1 no partial stores, 2 no MMAs, 4 no bias sums
};

// MN-major SWIZZLE_128B shared-memory matrix descriptor
__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (sm_100)
  d |= static_cast<uint64_t>(2) << 61;  // SWIZZLE_128B
  return d;
}
// kind::f16, D = f32, A = B = bf16, both MN-major
__host__ __device__ constexpr uint32_t make_idesc_bf16_f32_mn(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

__global__ void __launch_bounds__(kWtThreads, 1)
wgrad_c64_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                    WgradTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* xring = smem;
  uint8_t* dring = smem + kXS * kXSlotB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(dring + kDS * kDSlotB);
  uint64_t* xfull = bars;
  uint64_t* xempty = xfull + kXS;
  uint64_t* dfull = xempty + kXS;
  uint64_t* dempty = dfull + kDS;
  uint64_t* accfull = dempty + kDS;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(accfull + 1);
  float* red = reinterpret_cast<float*>(dring + kDS * kDSlotB + 256);  // [16 pixel lanes][64 co]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = a.H, nseg = a.nseg;
  const long long G = static_cast<long long>(a.B) * nseg * H;
  const int g0 = static_cast<int>(G * blockIdx.x / gridDim.x);
  const int g1 = static_cast<int>(G * (blockIdx.x + 1) / gridDim.x);
  const int Hp = H + 2;
  auto padded = [&](int g) { return (g / H) * Hp + (g % H) + 1; };
  const int pr_first = (g0 < g1) ? padded(g0) - 1 : 0;
  const int n_last = (g0 < g1) ? padded(g1 - 1) + 1 - pr_first : -1;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_dy);
    for (int i = 0; i < kXS; ++i) { mbar_init(&xfull[i], 1); mbar_init(&xempty[i], 1); }
    for (int i = 0; i < kDS; ++i) { mbar_init(&dfull[i], 1); mbar_init(&dempty[i], 5); }  // MMA commit + 4 bias warps
    mbar_init(accfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_holder);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  // programmatic dependent launch: the set-up above overlaps the predecessor's tail; the partial-sum reduction that
  // follows may become resident early (this grid is a single wave, so it cannot starve us) and waits for our exit
  if (threadIdx.x == 0) grid_dep_launch();
  grid_dep_wait();  // dY / X are produced by the preceding kernels

  if (g0 < g1) {
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (elect_one()) {
        int nx = 0;
        for (int g = g0, it = 0; g < g1; ++g, ++it) {
          const int nc = padded(g) - pr_first;
          const int upto = min(nc + 3, n_last);
          for (; nx <= upto; ++nx) {
            const int slot = nx % kXS;
            mbar_wait(&xempty[slot], ((nx / kXS) & 1) ^ 1, 21);
            const int pr = pr_first + nx;
            const int col = pr / Hp, yy = pr % Hp - 1;
            mbar_arrive_expect_tx(&xfull[slot], 130 * 128);
            tma_load_4d(xring + slot * kXSlotB, &tmap_x, &xfull[slot], 0, (col % nseg) * 128 - 1, yy, col / nseg);
          }
          const int ds = it % kDS;
          mbar_wait(&dempty[ds], ((it / kDS) & 1) ^ 1, 22);
          const int col = g / H, y = g % H;
          mbar_arrive_expect_tx(&dfull[ds], 128 * 128);
          tma_load_4d(dring + ds * kDSlotB, &tmap_dy, &dfull[ds], 0, (col % nseg) * 128, y, col / nseg);
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      const bool leader = elect_one();
      constexpr uint32_t id128 = make_idesc_bf16_f32_mn(128, 64);
      constexpr uint32_t id64 = make_idesc_bf16_f32_mn(64, 64);
      int released = 0;
      for (int g = g0, it = 0; g < g1; ++g, ++it) {
        const int nc = padded(g) - pr_first;
        const int nc_next = (g + 1 < g1) ? padded(g + 1) - pr_first : n_last + 2;
        const int ds = it % kDS;
        mbar_wait(&dfull[ds], (it / kDS) & 1, 23);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const int n = nc - 1 + d;
          mbar_wait(&xfull[n % kXS], (n / kXS) & 1, 24);
        }
        tcgen05_fence_after();
        const int seg = (g / H) % nseg;
        const int npx = min(128, a.W - seg * 128);
        const int ksteps = (npx + 15) >> 4;
        if (leader) {
          const uint32_t dbase = smem_u32(dring + ds * kDSlotB);
          for (int ks = 0; ks < ((a.probe & 2) ? 0 : ksteps); ++ks) {
            const uint64_t bdesc = make_sw128_mnmajor_desc(dbase + ks * 16 * 128, 1024, 1024);
            const uint32_t accum = (it > 0 || ks > 0) ? 1u : 0u;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const uint32_t xbase = smem_u32(xring + ((nc - 1 + dy) % kXS) * kXSlotB) + ks * 16 * 128;
              // rows (dx in {0,1}, ci): second MN atom = the row shifted by one pixel (LBO = 128 B)
              umma_f16_ss(tmem_base + dy * 64, make_sw128_mnmajor_desc(xbase, 128, 1024), bdesc, id128, accum);
              // rows ci of tap dx = 2
              umma_f16_ss(tmem_base + 192 + dy * 64, make_sw128_mnmajor_desc(xbase + 2 * 128, 1024, 1024), bdesc, id64,
                          accum);
            }
          }
          umma_commit(&dempty[ds]);
          for (int rel = released; rel <= nc_next - 2; ++rel) umma_commit(&xempty[rel % kXS]);
        }
        released = max(released, nc_next - 1);
        __syncwarp();
      }
      if (leader) umma_commit(accfull);
    } else {
      // ===================== bias gradient while the rows stream, then the epilogue =====================
      const int q = warp & 3;
      const int L = q * 32 + lane;  // TMEM lane
      {
        const int et = threadIdx.x - 64;     // 0..127
        const int ch = et & 7, pl = et >> 3;  // 8 channels (one 16-byte chunk) x 16 pixel lanes
        float acc8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int g = g0, it = 0; g < g1; ++g, ++it) {
          const int ds = it % kDS;
          mbar_wait(&dfull[ds], (it / kDS) & 1, 25);
          const int seg = (g / H) % nseg;
          const int npx = min(128, a.W - seg * 128);
          const uint8_t* base = dring + ds * kDSlotB;
          for (int p = pl; p < ((a.probe & 4) ? 0 : npx); p += 16) {
            const uint4 raw = *reinterpret_cast<const uint4*>(base + p * 128 + ((ch ^ (p & 7)) << 4));
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = __bfloat1622float2(h2[j]);
              acc8[2 * j] += f.x;
              acc8[2 * j + 1] += f.y;
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&dempty[ds]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) red[pl * 64 + ch * 8 + j] = acc8[j];
        named_bar_sync(1, 128);
        if (et < 64) {
          float t = 0.f;
#pragma unroll
          for (int k = 0; k < 16; ++k) t += red[k * 64 + et];
          a.dbpart[static_cast<size_t>(blockIdx.x) * 64 + et] = t;
        }
      }
      mbar_wait(accfull, 0, 26);
      tcgen05_fence_after();
      // partial layout [tap][co][ci]: for a fixed accumulator column (co) the 32 lanes of a warp hold 32 consecutive
      // ci, so every store instruction writes one full 128-byte line
      float* pbase = a.part + static_cast<size_t>(blockIdx.x) * 9 * 64 * 64;
#pragma unroll 1
      for (int dy = 0; dy < ((a.probe & 1) ? 0 : 3); ++dy) {
        {  // 128-lane accumulator: lane = dx*64 + ci (dx in {0,1}), column = co
          float* dst = pbase + static_cast<size_t>(dy * 3 + (L >> 6)) * 64 * 64 + (L & 63);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t rv[32];
            tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + dy * 64 + h * 32, rv);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) dst[(h * 32 + c) * 64] = __uint_as_float(rv[c]);
          }
        }
        {  // 64-row accumulator (M = 64): row r lives in lane (r % 16) + 32 * (r / 16); tap dx = 2
          float* dst = pbase + static_cast<size_t>(dy * 3 + 2) * 64 * 64 + (q * 16 + (lane & 15));
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t rv[32];
            tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + 192 + dy * 64 + h * 32, rv);
            tmem_ld_wait();
            if (lane < 16) {
#pragma unroll
              for (int c = 0; c < 32; ++c) dst[(h * 32 + c) * 64] = __uint_as_float(rv[c]);
            }
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

// Contract: partial sums + bias partials into scratch, reduced by wgrad_reduce.
int wgrad_c64_tc(const void* dy, long long dy_pix, long long dy_row, long long dy_img, const void* x, float* scratch, int B,
                 int H, int W, int num_sms, cudaStream_t s, int* S_out) {
  if (B <= 0 || H <= 0 || W <= 0) return DFIR_ERR_ARG;
  static bool configured[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return DFIR_ERR_CUDA;
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    if (cudaFuncSetAttribute(wgrad_c64_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWtSmem) != cudaSuccess)
      return DFIR_ERR_CUDA;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  const int grid = wgrad_c64_grid(B, H, W, num_sms);
  *S_out = grid;
  CUtensorMap tx, td;
  int rc = make_tmap_nhwc_bf16(&tx, x, 64, W, H, B, 128, static_cast<long long>(W) * 128,
                               static_cast<long long>(H) * W * 128, 130);
  if (rc != DFIR_OK) return rc;
  rc = make_tmap_nhwc_bf16(&td, dy, 64, W, H, B, dy_pix > 0 ? dy_pix : 128,
                           dy_row > 0 ? dy_row : static_cast<long long>(W) * 128,
                           dy_img > 0 ? dy_img : static_cast<long long>(H) * W * 128, 128);
  if (rc != DFIR_OK) return rc;
  WgradTcArgs a{};
  a.part = scratch;
  a.dbpart = scratch + static_cast<size_t>(grid) * 9 * 64 * 64;
  a.B = B; a.H = H; a.W = W; a.nseg = (W + 127) / 128;
#ifdef DFIR_PROBES  // timing experiments (wrong results) exist only in a `make PROBES=1` build
  a.probe = getenv("DFIR_WGRAD_PROBE") != nullptr ? atoi(getenv("DFIR_WGRAD_PROBE")) : 0;
#else
  a.probe = 0;
#endif
  return launch_pdl(PDL_WGRAD, wgrad_c64_tc_kernel, dim3(grid), dim3(kWtThreads), kWtSmem, s, tx, td, a) == cudaSuccess
             ? DFIR_OK
             : DFIR_ERR_CUDA;
}

int wgrad_tc_watchdog(unsigned int* out8, int reset) {
  if (cudaMemcpyFromSymbol(out8, ptx::g_dfir_watchdog, 8 * sizeof(unsigned int)) != cudaSuccess) return DFIR_ERR_CUDA;
  if (reset) {
    unsigned int z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (cudaMemcpyToSymbol(ptx::g_dfir_watchdog, z, sizeof(z)) != cudaSuccess) return DFIR_ERR_CUDA;
  }
  return DFIR_OK;
}

// number of partial-sum CTAs (= scratch slices) of a weight-gradient launch
int wgrad_c64_grid(int B, int H, int W, int num_sms) {
  const long long G = static_cast<long long>(B) * ((W + 127) / 128) * H;
  int grid = num_sms > 0 ? num_sms : 148;
  if (G < grid) grid = static_cast<int>(G);
  return grid < 1 ? 1 : grid;
}

int wgrad_c64_co_major() { return 1; }

int wgrad_c64(const void* dy, long long dy_pix, long long dy_row, long long dy_img, const void* x, float* scratch, int B,
              int H, int W, int num_sms, cudaStream_t s, int* S_out) {
  return wgrad_c64_tc(dy, dy_pix, dy_row, dy_img, x, scratch, B, H, W, num_sms, s, S_out);
}

}  // namespace dfir
