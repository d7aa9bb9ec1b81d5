// Weight gradient of the 64 -> 64 3x3 convolution on the tensor cores (bf16 operands, fp32 accumulate):
//     dW[co][ci][dy][dx] = sum over pixels p of dY[p][co] * X[p + (dy-1, dx-1)][ci]      (zero padded)
//     db[co]             = sum over pixels p of dY[p][co]
// i.e. the backward of default_conv (/root/reference/Code/SISR/models/advanced/common.py:5-8) with respect to its
// parameters, as autograd computes it inside BaseModel.standard_update (models/__init__.py:481-489).
//
// GEMM view per tap: D[64 co][64 ci] += A[co][K = pixels] * B[K][ci].  Both operands are NHWC, i.e. the reduction
// dimension (pixels) is the OUTER dimension of both — "MN-major" operands.  This first version uses warp-level
// mma.sync.m16n8k16 (ldmatrix.trans turns the pixel-major shared-memory rows into K-fragments): nine warps, one per
// tap, all reading the same dY row and their own shifted view of the X rows.  A persistent CTA owns a band of image
// rows, keeps the 9 x 64 x 64 fp32 partial sums in registers, and writes them once; a second kernel reduces over
// CTAs in a fixed order (deterministic) and stores OIHW.  (The tcgen05 form needs MN-major shared-memory
// descriptors; DESIGN.md lists it as the next step for this kernel.)
//
// Shared memory: ring of 5 X rows (130 px incl. the x halo, 128-byte rows, 16-byte chunks XOR-swizzled by the row
// index so that ldmatrix is bank-conflict free) + 3 dY row buffers, filled by cp.async two rows ahead.
#include "kernels.h"
#include "ptx.cuh"
#include "launch.cuh"

namespace dfir {

namespace {

constexpr int kWgThreads = 384;  // 9 MMA warps (one per tap) + 3 helper warps (bias gradient); registers are
                                 // allocated per 4 warps, so 9 warps would not get more registers than 12
constexpr int kXSlotBytes = 130 * 128;
constexpr int kDyBytes = 128 * 128;
constexpr int kXSlots = 5;   // 3 live rows + 2 in flight
constexpr int kDyBufs = 3;   // current + 2 in flight
constexpr int kWgSmem = kXSlots * kXSlotBytes + kDyBufs * kDyBytes;

struct WgradArgs {
  const uint8_t* dy;  // bf16, byte strides below
  long long dy_pix, dy_row, dy_img;
  const __nv_bfloat16* x;  // dense NHWC, 64 channels
  float* part;             // [grid][9][64 ci][64 co]
  float* dbpart;           // [grid][64]
  int B, H, W, nseg;
};

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ptx::smem_u32(dst)), "l"(src) : "memory");
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_c64_mma_kernel(WgradArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  ptx::grid_dep_wait();
  ptx::grid_dep_launch();
  uint8_t* xring = smem;
  uint8_t* dybuf = smem + kXSlots * kXSlotBytes;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = a.H, W = a.W, nseg = a.nseg;
  const long long G = static_cast<long long>(a.B) * nseg * H;
  const int g0 = static_cast<int>(G * blockIdx.x / gridDim.x);
  const int g1 = static_cast<int>(G * (blockIdx.x + 1) / gridDim.x);
  const int tdy = warp / 3, tdx = warp % 3;

  float acc[4][8][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;
  float dbacc = 0.f;  // helper warps: bias gradient of channel (tid - 288) & 63

  auto load_x_row = [&](int b, int yy, int seg) {
    uint8_t* slot = xring + ((yy + 1) % kXSlots) * kXSlotBytes;
    const bool row_ok = yy >= 0 && yy < H;
    for (int i = tid; i < 130 * 8; i += kWgThreads) {
      const int p = i >> 3, ch = i & 7;
      const int x = seg * 128 + p - 1;
      uint8_t* dst = slot + p * 128 + ((ch ^ (p & 7)) << 4);
      if (row_ok && x >= 0 && x < W)
        cp_async16(dst, a.x + ((static_cast<size_t>(b) * H + yy) * W + x) * 64 + ch * 8);
      else
        *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
    }
  };
  auto load_dy_row = [&](int b, int y, int seg, int buf) {
    uint8_t* dst0 = dybuf + buf * kDyBytes;
    const uint8_t* src0 = a.dy + static_cast<long long>(b) * a.dy_img + static_cast<long long>(y) * a.dy_row;
    for (int i = tid; i < 128 * 8; i += kWgThreads) {
      const int p = i >> 3, ch = i & 7;
      const int x = seg * 128 + p;
      uint8_t* dst = dst0 + p * 128 + ((ch ^ (p & 7)) << 4);
      if (x < W)
        cp_async16(dst, src0 + static_cast<long long>(x) * a.dy_pix + ch * 16);
      else
        *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
    }
  };

  for (int g = g0, it = 0; g < g1; ++g, ++it) {
    const int col = g / H, y = g % H;
    const int b = col / nseg, seg = col % nseg;
    if (it == 0 || y == 0) {  // start of a (image, segment) column: nothing is in flight
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      load_x_row(b, y - 1, seg);
      load_x_row(b, y, seg);
      load_x_row(b, y + 1, seg);
      load_dy_row(b, y, seg, it % kDyBufs);
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (g + 1 < g1 && y + 1 < H) {
        load_x_row(b, y + 2, seg);
        load_dy_row(b, y + 1, seg, (it + 1) % kDyBufs);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.wait_group 1;" ::: "memory");  // everything but the newest group: this row is complete
    __syncthreads();
    if (g + 2 < g1 && y + 2 < H) {  // two rows ahead: one new X row + its dY row
      load_x_row(b, y + 3, seg);
      load_dy_row(b, y + 2, seg, (it + 2) % kDyBufs);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    const int npx = min(128, W - seg * 128);
    const int ksteps = (npx + 15) >> 4;
    const uint32_t dy_s = ptx::smem_u32(dybuf + (it % kDyBufs) * kDyBytes);
    const uint32_t x_s = ptx::smem_u32(xring + ((y + tdy) % kXSlots) * kXSlotBytes);  // row y + tdy - 1
    const int q = lane >> 3, r8 = lane & 7;
    if (warp >= 9) {
      // bias gradient: column sums of the dY row straight from shared memory (threads 288..351, one channel each)
      const int co = tid - 288;
      if (co < 64) {
        const uint8_t* base = dybuf + (it % kDyBufs) * kDyBytes + (co & 7) * 2;
        const int ch = co >> 3;
        for (int p = 0; p < npx; ++p)
          dbacc += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(base + p * 128 + ((ch ^ (p & 7)) << 4)));
      }
      continue;
    }
    for (int ks = 0; ks < ksteps; ++ks) {
      const int k0 = ks * 16;
      uint32_t af[4][4];
      {
        const int krow = k0 + (q >> 1) * 8 + r8;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
          const int chunk = mt * 2 + (q & 1);
          ldmatrix_x4_trans(dy_s + krow * 128 + ((chunk ^ (krow & 7)) << 4), af[mt][0], af[mt][1], af[mt][2], af[mt][3]);
        }
      }
      const int prow = k0 + (q & 1) * 8 + r8 + tdx;  // ring pixel index = x_local + dx
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b0a, b1a, b0b, b1b;
        const int chunk = np * 2 + (q >> 1);
        ldmatrix_x4_trans(x_s + prow * 128 + ((chunk ^ (prow & 7)) << 4), b0a, b1a, b0b, b1b);
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
          mma_bf16_16816(acc[mt][2 * np], af[mt], b0a, b1a);
          mma_bf16_16816(acc[mt][2 * np + 1], af[mt], b0b, b1b);
        }
      }
    }
  }

  // partial sums of this CTA: part[cta][tap][ci][co]
  if (warp >= 9) {
    if (tid - 288 < 64) a.dbpart[static_cast<size_t>(blockIdx.x) * 64 + (tid - 288)] = dbacc;
  } else {
    const int gq = lane >> 2, t4 = lane & 3;
    float* p = a.part + (static_cast<size_t>(blockIdx.x) * 9 + warp) * 64 * 64;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int co = mt * 16 + gq, ci = nt * 8 + 2 * t4;
        p[ci * 64 + co] = acc[mt][nt][0];
        p[(ci + 1) * 64 + co] = acc[mt][nt][1];
        p[ci * 64 + co + 8] = acc[mt][nt][2];
        p[(ci + 1) * 64 + co + 8] = acc[mt][nt][3];
      }
  }
}

}  // namespace

int wgrad_c64_grid(int B, int H, int W, int num_sms) {
  const long long G = static_cast<long long>(B) * ((W + 127) / 128) * H;
  int grid = num_sms > 0 ? num_sms : 148;
  if (G < grid) grid = static_cast<int>(G);
  return grid < 1 ? 1 : grid;
}

// scratch: grid * (9*64*64 + 64) floats.  Launches the partial-sum kernel only; the caller reduces with wgrad_reduce.
int wgrad_c64_bf16(const void* dy, long long dy_pix, long long dy_row, long long dy_img, const void* x, float* scratch,
                   int B, int H, int W, int num_sms, cudaStream_t s, int* S_out) {
  if (B <= 0 || H <= 0 || W <= 0) return DFIR_ERR_ARG;
  static bool configured[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return DFIR_ERR_CUDA;
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    if (cudaFuncSetAttribute(wgrad_c64_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem) != cudaSuccess)
      return DFIR_ERR_CUDA;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  const int grid = wgrad_c64_grid(B, H, W, num_sms);
  *S_out = grid;
  WgradArgs a{};
  a.dy = reinterpret_cast<const uint8_t*>(dy);
  a.dy_pix = dy_pix > 0 ? dy_pix : 128;
  a.dy_row = dy_row > 0 ? dy_row : static_cast<long long>(W) * 128;
  a.dy_img = dy_img > 0 ? dy_img : static_cast<long long>(H) * W * 128;
  a.x = reinterpret_cast<const __nv_bfloat16*>(x);
  a.part = scratch;
  a.dbpart = scratch + static_cast<size_t>(grid) * 9 * 64 * 64;
  a.B = B; a.H = H; a.W = W; a.nseg = (W + 127) / 128;
  return launch_pdl(PDL_SIMT, wgrad_c64_mma_kernel, dim3(grid), dim3(kWgThreads), kWgSmem, s, a) == cudaSuccess ? DFIR_OK : DFIR_ERR_CUDA;
}

}  // namespace dfir
