// Micro-benchmark: issue rate of tcgen05.mma kind::f16 for the conv's operand shapes on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_probe tools/mma_probe.cu -lcuda
// Variants: A from tensor memory (TS) or shared memory (SS); N = 64 / 128 / 256; cta_group::1 (M = 128) or
// cta_group::2 (M = 256 over a CTA pair, each CTA holding half of the B tile); optional concurrent "loader" warps that
// run the conv kernel's smem -> TMEM copy pattern (LDS.128 + tcgen05.st) once per batch or flat out.
// The numbers decide whether cta_group::2 is worth building into conv_tc.cu (profiles/r02_mma_probe.md).
#include "../super-resolution-meta-attention-networks_b200/csrc/ptx.cuh"

#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace dfir::ptx;

namespace {

constexpr int kBatch = 36;  // MMAs per commit, as one conv row

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_alloc_cg(uint32_t* dst, uint32_t cols) {
  if constexpr (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc_cg(uint32_t addr, uint32_t cols) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
template <int CG, bool TS>
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  if constexpr (CG == 1 && TS)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
  else if constexpr (CG == 2 && TS)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
  else if constexpr (CG == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
template <int CG>
__device__ __forceinline__ void commit(uint64_t* bar) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3)) : "memory");
}
__device__ __forceinline__ void spin_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// threads: warp 0 = issuer, warps 1..4 = loaders (optional), total 160
template <int CG, bool TS, int N>
__global__ void __launch_bounds__(160, 1) probe_kernel(int iters, int load_mode, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int NB = N / CG;               // B rows held by this CTA
  constexpr int kTapBytes = NB * 128;      // one tap (64 k-elements) of B
  constexpr int kTaps = (9 * kTapBytes <= 147456) ? 9 : (147456 / kTapBytes);
  uint8_t* bsm = smem;                     // kTaps tap tiles
  uint8_t* asm_ = smem + 147456;           // 16 KB A tile (SS) or loader source row (17 KB)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 147456 + 3 * 17408);
  uint32_t* holder = reinterpret_cast<uint32_t*>(bars + 4);
  volatile int* issued = reinterpret_cast<volatile int*>(holder + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;

  for (int i = threadIdx.x; i < (147456 + 3 * 17408) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = (i & 1) ? 0x3a833a83u : 0x3f803a83u;  // bf16 1.0 / 0.001
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    *issued = 0;
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 0) tmem_alloc_cg<CG>(holder, 512);
  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *holder;
  // fill the A region of tensor memory (columns 256..511) with bf16 ones
  if (warp >= 1) {
    const int q = warp & 3;
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0x3f803f80u;
    for (int c = 256; c < 512; c += 32) tmem_st_32x32b_x32(tmem + (static_cast<uint32_t>(q * 32) << 16) + c, v);
    tmem_st_wait();
  }
  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tcgen05_fence_after();

  if (warp == 0) {
    if (rank == 0) {
      const bool leader = elect_one();
      constexpr uint32_t idesc = make_idesc_bf16_f32(128 * CG, N);
      const uint64_t db_base = make_sw128_kmajor_desc(smem_u32(bsm), 1024, 0);
      const uint64_t da_base = make_sw128_kmajor_desc(smem_u32(asm_), 1024, 0);
      long long t0 = 0, t1 = 0;
      unsigned long long g0 = 0, g1 = 0;
      constexpr int kAccs = (2 * N <= 256) ? 2 : 1;
      if (leader) {
        t0 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
        for (int it = 0; it < iters; ++it) {
          const uint32_t d = tmem + (it % kAccs) * N;
#pragma unroll
          for (int j = 0; j < kBatch; ++j) {
            const int tap = (j >> 2) % kTaps, k = j & 3;
            const uint64_t db = db_base + static_cast<uint32_t>((tap * kTapBytes + k * 32) >> 4);
            mma<CG, TS>(d, tmem + 256 + (j % 12) * 8, da_base + static_cast<uint32_t>((k * 32) >> 4), db, idesc, j != 0);
          }
          commit<CG>(&bars[it & 1]);
          *issued = it + 1;
          if (it >= 1) spin_wait(&bars[(it - 1) & 1], ((it - 1) >> 1) & 1);
        }
        spin_wait(&bars[(iters - 1) & 1], ((iters - 1) >> 1) & 1);
        t1 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        *issued = 1 << 30;
        out[blockIdx.x * 2] = static_cast<unsigned long long>(t1 - t0);
        out[blockIdx.x * 2 + 1] = g1 - g0;
      }
      __syncwarp();
    } else {
      // peer CTA of the pair: its barriers receive the multicast commits; wait for the last one
      if (lane == 0) {
        for (int it = 0; it < iters; ++it) {
          spin_wait(&bars[it & 1], (it >> 1) & 1);
          *issued = it + 1;
        }
        *issued = 1 << 30;
      }
      __syncwarp();
    }
  } else if (load_mode >= 3) {
    // epilogue pattern: tcgen05.ld of 64 accumulator columns (the accumulator the MMAs are NOT writing), once per
    // batch (mode 3) or flat out (mode 4); lane 0 of warp 1 reports clocks per 32-column load
    const int q = warp & 3;
    int done = 0;
    long long tl = 0;
    int nld = 0;
    while (true) {
      const int target = *issued;
      if (target >= (1 << 30)) break;
      if (load_mode == 3 && done >= target) continue;
      uint32_t v[32];
      const long long c0 = clock64();
      tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(q * 32) << 16) + ((done + 1) & 1) * N, v);
      tmem_ld_wait();
      tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(q * 32) << 16) + ((done + 1) & 1) * N + 32, v);
      tmem_ld_wait();
      tl += clock64() - c0;
      nld += 2;
      if (v[0] == 0x12345678u) out[0] = 1;  // keep the loads alive
      ++done;
    }
    if (warp == 1 && lane == 0 && rank == 0) out[296 + blockIdx.x] = nld > 0 ? static_cast<unsigned long long>(tl / nld) : 0ull;
  } else if (load_mode != 0) {
    // loader pattern of conv_tc.cu: one ring row -> three dx-shifted TMEM copies (into columns 352..447, unused by MMAs)
    const int q = warp & 3;
    const int m = q * 32 + lane;
    int done = 0;
    while (true) {
      const int target = *issued;
      if (target >= (1 << 30)) break;
      if (load_mode == 1 && done >= target) continue;
      const uint8_t* srow = asm_ + (done % 3) * 17408;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int p = m + dx;
        const uint4* src = reinterpret_cast<const uint4*>(srow + p * 128);
        uint32_t v[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 t = src[c ^ (p & 7)];
          v[4 * c + 0] = t.x; v[4 * c + 1] = t.y; v[4 * c + 2] = t.z; v[4 * c + 3] = t.w;
        }
        tmem_st_32x32b_x32(tmem + (static_cast<uint32_t>(q * 32) << 16) + 352 + dx * 32, v);
      }
      tmem_st_wait();
      ++done;
    }
  }
  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc_cg<CG>(tmem, 512);
  }
}

template <int CG, bool TS, int N>
void run(const char* name, int grid, int iters, int load_mode, unsigned long long* dout) {
  auto kern = probe_kernel<CG, TS, N>;
  const int smem = 147456 + 3 * 17408 + 64 + 1024;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(160);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaMemset(dout, 0, 148 * 24);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    cudaError_t rc = cudaLaunchKernelEx(&cfg, kern, iters, load_mode, dout);
    cudaEventRecord(e1);
    cudaError_t rs = cudaDeviceSynchronize();
    if (rc != cudaSuccess || rs != cudaSuccess) {
      printf("%-34s grid %3d FAILED: %s / %s\n", name, grid, cudaGetErrorString(rc), cudaGetErrorString(rs));
      return;
    }
  }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<unsigned long long> h(grid * 2);
  cudaMemcpy(h.data(), dout, grid * 16, cudaMemcpyDeviceToHost);
  unsigned long long ldclk = 0;
  cudaMemcpy(&ldclk, dout + 296, 8, cudaMemcpyDeviceToHost);
  double clk = 0, ns = 0;
  int n = 0;
  for (int b = 0; b < grid; b += CG) {
    clk += static_cast<double>(h[b * 2]);
    ns += static_cast<double>(h[b * 2 + 1]);
    ++n;
  }
  clk /= n;
  ns /= n;
  const double mmas = static_cast<double>(iters) * kBatch;
  const double flop = mmas * 2.0 * 128 * CG * N * 16 * n;
  const double floor_clk = 128.0 * N / 256.0;  // per CTA: M = 128 rows x N columns x K = 16 at 4096 MAC/clk/SM
  printf("%-34s grid %3d load %d : %7.1f clk/MMA (floor %5.1f, %5.1f %% of floor rate) %7.1f ns/MMA  SM clock %4.0f MHz  %7.1f TFLOP/s chip",
         name, grid, load_mode, clk / mmas, floor_clk, 100.0 * floor_clk / (clk / mmas), ns / mmas, clk / ns * 1000.0,
         flop / (ms * 1e-3) * 1e-12);
  if (load_mode >= 3) printf("  tcgen05.ld.32x32b.x32 + wait: %llu clk", ldclk);
  printf("\n");
  fflush(stdout);
}

}  // namespace

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  unsigned long long* dout = nullptr;
  cudaMalloc(&dout, 148 * 16 + 148 * 8);
  const int only_ld = argc > 2 ? atoi(argv[2]) : 0;
  if (only_ld == 2) {  // issue cost: tiny N keeps the tensor pipe nearly idle, so clk/MMA = cost of issuing one UTCHMMA
    run<1, true, 8>("cg1 TS N=8 (issue cost)", 148, iters, 0, dout);
    run<1, true, 16>("cg1 TS N=16", 148, iters, 0, dout);
    run<1, true, 32>("cg1 TS N=32", 148, iters, 0, dout);
    run<1, true, 64>("cg1 TS N=64", 148, iters, 0, dout);
    run<1, true, 192>("cg1 TS N=192", 148, iters, 0, dout);
    return 0;
  }
  if (only_ld) {
    for (int lm : {0, 3, 4}) {
      run<1, true, 64>("cg1 TS N=64", 148, iters, lm, dout);
      run<1, true, 128>("cg1 TS N=128", 148, iters, lm, dout);
      run<2, true, 64>("cg2 TS N=64 (M=256)", 148, iters, lm, dout);
    }
    return 0;
  }
  for (int grid : {8, 148}) {
    for (int lm : {0, 1, 2}) {
      run<1, true, 64>("cg1 TS N=64", grid, iters, lm, dout);
      run<1, true, 128>("cg1 TS N=128", grid, iters, lm, dout);
      run<1, true, 256>("cg1 TS N=256", grid, iters, lm, dout);
      run<1, false, 64>("cg1 SS N=64", grid, iters, lm, dout);
      run<1, false, 128>("cg1 SS N=128", grid, iters, lm, dout);
      run<1, false, 256>("cg1 SS N=256", grid, iters, lm, dout);
      run<2, true, 64>("cg2 TS N=64 (M=256)", grid, iters, lm, dout);
      run<2, true, 128>("cg2 TS N=128 (M=256)", grid, iters, lm, dout);
      run<2, true, 256>("cg2 TS N=256 (M=256)", grid, iters, lm, dout);
      run<2, false, 64>("cg2 SS N=64 (M=256)", grid, iters, lm, dout);
      run<2, false, 256>("cg2 SS N=256 (M=256)", grid, iters, lm, dout);
    }
  }
  return 0;
}
