// Online degradation of the training input pipeline on the GPU:
//   BatchBlur        (reference Code/sr_tools/gaussian_utils.py:346-368): reflection pad l/2, one l x l kernel per image
//                    (or one shared kernel) applied to every colour plane as a cross-correlation;
//   PCAEncoder       (:333-343): kernel code = flattened kernel (l*l) times the PCA matrix (l*l x k);
//   SRMDPreprocessing (:371-424): blur -> + sigma_b * noise, clamp to [0,1] -> code = [kernel code, 10 sigma_b].
// One kernel does blur + noise + clamp: each CTA computes a 32 x 32 output tile of one (image, plane) from a reflected
// (32 + l - 1)^2 input tile in shared memory; a thread computes 4 outputs of a row from a sliding window of l + 3 inputs
// in registers (l + 3 + l shared-memory reads per 4 l FMAs).  4 B in + 4 B out per pixel; l^2 = 441 MACs per pixel at l = 21.
#include "kernels.h"

#include <algorithm>

namespace dfir {

namespace {

constexpr int kBlurTile = 32;
constexpr int kBlurMaxL = 33;

__device__ __forceinline__ int reflect(int i, int n) {  // nn.ReflectionPad2d index (no edge repeat), |pad| < n
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

template <int L>  // L = kernel size known at compile time (window and loops in registers), 0 = any size up to kBlurMaxL
__global__ void __launch_bounds__(256)
blur_noise_kernel(const float* __restrict__ x, const float* __restrict__ kernels, int kernel_per_image,
                  const float* __restrict__ noise, const float* __restrict__ sigma, float* __restrict__ out, int C, int H,
                  int W, int l_dyn, int pad_lo, int clamp01) {
  const int l = L > 0 ? L : l_dyn;
  extern __shared__ float sm[];
  const int tw = kBlurTile + l - 1;          // input tile edge
  float* tile = sm;                           // [tw][tw + 1]
  float* kk = sm + tw * (tw + 1);             // [l][l]
  const int plane = blockIdx.z;               // b * C + c
  const int b = plane / C;
  const int y0 = blockIdx.y * kBlurTile, x0 = blockIdx.x * kBlurTile;
  const float* xp = x + static_cast<size_t>(plane) * H * W;
  const float* kp = kernels + (kernel_per_image ? static_cast<size_t>(b) * l * l : 0);
  for (int i = threadIdx.x; i < l * l; i += blockDim.x) kk[i] = kp[i];
  for (int i = threadIdx.x; i < tw * tw; i += blockDim.x) {
    const int ty = i / tw, tx = i % tw;
    const int yy = reflect(y0 + ty - pad_lo, H), xx = reflect(x0 + tx - pad_lo, W);
    tile[ty * (tw + 1) + tx] = xp[static_cast<size_t>(yy) * W + xx];
  }
  __syncthreads();
  const int oy = threadIdx.x >> 3, ox = (threadIdx.x & 7) * 4;  // 32 rows x 8 groups of 4 outputs
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int ky = 0; ky < l; ++ky) {
    const float* trow = tile + (oy + ky) * (tw + 1) + ox;
    const float* krow = kk + ky * l;
    float win[(L > 0 ? L : kBlurMaxL) + 3];
#pragma unroll
    for (int j = 0; j < (L > 0 ? L : kBlurMaxL) + 3; ++j)
      if (L > 0 || j < l + 3) win[j] = trow[j];
#pragma unroll
    for (int kx = 0; kx < (L > 0 ? L : kBlurMaxL); ++kx) {
      if (L == 0 && kx >= l) break;
      const float w = krow[kx];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaf(win[kx + j], w, acc[j]);
    }
  }
  const int y = y0 + oy;
  if (y >= H) return;
  const float sg = (noise != nullptr && sigma != nullptr) ? sigma[b] : 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int xo = x0 + ox + j;
    if (xo >= W) break;
    const size_t e = static_cast<size_t>(plane) * H * W + static_cast<size_t>(y) * W + xo;
    float v = acc[j];
    if (noise != nullptr && sigma != nullptr) v = fmaf(noise[e], sg, v);
    if (clamp01) v = fminf(fmaxf(v, 0.f), 1.f);
    out[e] = v;
  }
}

// code[b][j] = sum_i kernel[b][i] * pca[i][j] (+ code[b][k] = 10 sigma[b]); one CTA per image, fixed summation order
__global__ void __launch_bounds__(128)
pca_encode_kernel(const float* __restrict__ kernels, const float* __restrict__ pca, const float* __restrict__ sigma,
                  float* __restrict__ code, int n, int k, int code_stride) {
  __shared__ float part[4][32];
  const int b = blockIdx.x;
  const int j = threadIdx.x & 31, slice = threadIdx.x >> 5;
  float t = 0.f;
  if (j < k)
    for (int i = slice; i < n; i += 4) t = fmaf(kernels[static_cast<size_t>(b) * n + i], pca[static_cast<size_t>(i) * k + j], t);
  part[slice][j] = t;
  __syncthreads();
  if (threadIdx.x < k) code[static_cast<size_t>(b) * code_stride + threadIdx.x] =
      (part[0][threadIdx.x] + part[1][threadIdx.x]) + (part[2][threadIdx.x] + part[3][threadIdx.x]);
  if (threadIdx.x == 0 && sigma != nullptr) code[static_cast<size_t>(b) * code_stride + k] = 10.f * sigma[b];
}

}  // namespace

int batch_blur(const float* x, const float* kernels, int kernel_per_image, const float* noise, const float* sigma, float* out,
               int B, int C, int H, int W, int l, int clamp01, cudaStream_t s) {
  if (l < 1 || l > kBlurMaxL || l / 2 >= H || l / 2 >= W) return DFIR_ERR_ARG;  // reflection needs pad < size
  if (B <= 0 || C <= 0) return DFIR_OK;
  const int tw = kBlurTile + l - 1;
  const int smem = (tw * (tw + 1) + l * l) * 4;
  if (smem > 48 * 1024) return DFIR_ERR_ARG;
  if (static_cast<long long>(B) * C > 65535) return DFIR_ERR_ARG;
  dim3 grid((W + kBlurTile - 1) / kBlurTile, (H + kBlurTile - 1) / kBlurTile, B * C);
  if (l == 21)
    blur_noise_kernel<21><<<grid, 256, smem, s>>>(x, kernels, kernel_per_image, noise, sigma, out, C, H, W, l, l / 2, clamp01);
  else if (l == 15)
    blur_noise_kernel<15><<<grid, 256, smem, s>>>(x, kernels, kernel_per_image, noise, sigma, out, C, H, W, l, l / 2, clamp01);
  else
    blur_noise_kernel<0><<<grid, 256, smem, s>>>(x, kernels, kernel_per_image, noise, sigma, out, C, H, W, l, l / 2, clamp01);
  return cudaGetLastError() == cudaSuccess ? DFIR_OK : DFIR_ERR_CUDA;
}

int pca_encode(const float* kernels, const float* pca, const float* sigma, float* code, int B, int n, int k, cudaStream_t s) {
  if (k < 1 || k > 32 || n < 1) return DFIR_ERR_ARG;
  if (B <= 0) return DFIR_OK;
  pca_encode_kernel<<<B, 128, 0, s>>>(kernels, pca, sigma, code, n, k, k + (sigma != nullptr ? 1 : 0));
  return cudaGetLastError() == cudaSuccess ? DFIR_OK : DFIR_ERR_CUDA;
}

}  // namespace dfir
