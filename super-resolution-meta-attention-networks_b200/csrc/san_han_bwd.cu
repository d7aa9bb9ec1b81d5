// Backward kernels of the Q-HAN / Q-SAN specific layers (fp32, NHWC; all HBM / latency bound):
//   region non-local attention     (autograd of advanced/SAN_blocks.py:104-148, 314-336)
//   LAM  layer attention           (autograd of advanced/HAN_blocks.py:24-37)
//   CSAM channel-spatial attention (autograd of advanced/HAN_blocks.py:59-76)
//   SOCA tail: column mean + FC-ReLU-FC-sigmoid (advanced/SAN_blocks.py:290-302) forward and backward
//   two generic reductions: per-image channel dot products and sum_p A[p] (x) B[p] (1x1-conv weight gradients)
// The reference has no hand-written backward for these layers (only Covpool / Sqrtm, see san_han.cu): what is restated
// here is what autograd derives from the forward.  Every reduction runs in a fixed two-stage order.
#include "kernels.h"

#include <algorithm>

namespace dfir {

namespace {

inline int ok_or_cuda3() { return cudaGetLastError() == cudaSuccess ? DFIR_OK : DFIR_ERR_CUDA; }

#define DFIR_TRY_RC(expr)             \
  do {                                \
    int rc__ = (expr);                \
    if (rc__ != DFIR_OK) return rc__; \
  } while (0)

constexpr int kDotChunks = 32;

// ------------------------------------------------------------------------------------------------
// channel_dot: out[b][c] = sum_p a[b][p][c] * bb[b][p][c]   (C a multiple of 4, C <= 256)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
channel_dot_partial_kernel(const float* __restrict__ a, const float* __restrict__ bb, float* __restrict__ partial,
                           long long HW, int C) {
  extern __shared__ float red[];  // [lanes][C]
  const int C4 = C / 4;
  const int lanes = 256 / C4;  // pixel lanes per block
  const int b = blockIdx.y, chunk = blockIdx.x;
  const long long per = (HW + kDotChunks - 1) / kDotChunks;
  const long long p0 = chunk * per, p1 = min(HW, p0 + per);
  const int c4 = threadIdx.x % C4, lane = threadIdx.x / C4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lane < lanes) {
    const float4* pa = reinterpret_cast<const float4*>(a + static_cast<size_t>(b) * HW * C);
    const float4* pb = reinterpret_cast<const float4*>(bb + static_cast<size_t>(b) * HW * C);
    for (long long p = p0 + lane; p < p1; p += lanes) {
      const float4 u = pa[p * C4 + c4], v = pb[p * C4 + c4];
      acc.x = fmaf(u.x, v.x, acc.x); acc.y = fmaf(u.y, v.y, acc.y);
      acc.z = fmaf(u.z, v.z, acc.z); acc.w = fmaf(u.w, v.w, acc.w);
    }
    float* r = red + static_cast<size_t>(lane) * C + c4 * 4;
    r[0] = acc.x; r[1] = acc.y; r[2] = acc.z; r[3] = acc.w;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float t = 0.f;
    for (int l = 0; l < lanes; ++l) t += red[static_cast<size_t>(l) * C + c];
    partial[(static_cast<size_t>(b) * kDotChunks + chunk) * C + c] = t;
  }
}

// out[b][c] = sum over chunks; total (optional) = sum over (b, c) in index order (+ what it held when accumulate != 0)
__global__ void channel_dot_final_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                         float* __restrict__ total, int BC, int C, int accumulate) {
  __shared__ float sums[256];
  float mine = 0.f;
  for (int i = threadIdx.x; i < BC; i += 256) {
    const int b = i / C, c = i % C;
    float t = 0.f;
    for (int k = 0; k < kDotChunks; ++k) t += partial[(static_cast<size_t>(b) * kDotChunks + k) * C + c];
    if (out != nullptr) out[i] = t;
    mine += t;
  }
  sums[threadIdx.x] = mine;
  __syncthreads();
  if (threadIdx.x == 0 && total != nullptr) {
    float t = 0.f;
    for (int i = 0; i < 256; ++i) t += sums[i];
    *total = accumulate ? *total + t : t;
  }
}

// ------------------------------------------------------------------------------------------------
// outer_reduce: out[o][i] = sum_p A[p][o] * Bm[p][i], colsum[o] = sum_p A[p][o]   (Ao * (Bi + 1) <= 2048)
// ------------------------------------------------------------------------------------------------
constexpr int kOuterTile = 32;   // pixels staged per step
constexpr int kOuterMaxPer = 8;  // outputs per thread

__global__ void __launch_bounds__(256)
outer_reduce_partial_kernel(const float* __restrict__ A, int Ao, const float* __restrict__ Bm, int Bi, long long npix,
                            float* __restrict__ partial, int nchunk) {
  extern __shared__ float sm[];  // A tile [32][Ao], B tile [32][Bi + 1] (last column = 1)
  float* sa = sm;
  float* sb = sm + kOuterTile * Ao;
  const int Bi1 = Bi + 1;
  const int nout = Ao * Bi1;
  const long long per = (npix + nchunk - 1) / nchunk;
  const long long p0 = blockIdx.x * per, p1 = min(npix, p0 + per);
  float acc[kOuterMaxPer];
#pragma unroll
  for (int k = 0; k < kOuterMaxPer; ++k) acc[k] = 0.f;
  for (long long t0 = p0; t0 < p1; t0 += kOuterTile) {
    const int np = static_cast<int>(min(static_cast<long long>(kOuterTile), p1 - t0));
    __syncthreads();
    for (int i = threadIdx.x; i < kOuterTile * Ao; i += 256) sa[i] = i < np * Ao ? A[t0 * Ao + i] : 0.f;
    for (int i = threadIdx.x; i < kOuterTile * Bi1; i += 256) {
      const int p = i / Bi1, c = i % Bi1;
      sb[i] = p < np ? (c < Bi ? Bm[(t0 + p) * Bi + c] : 1.f) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kOuterMaxPer; ++k) {
      const int idx = threadIdx.x + k * 256;
      if (idx < nout) {
        const int o = idx / Bi1, i = idx % Bi1;
        float t = acc[k];
        for (int p = 0; p < kOuterTile; ++p) t = fmaf(sa[p * Ao + o], sb[p * Bi1 + i], t);
        acc[k] = t;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kOuterMaxPer; ++k) {
    const int idx = threadIdx.x + k * 256;
    if (idx < nout) partial[static_cast<size_t>(blockIdx.x) * nout + idx] = acc[k];
  }
}

__global__ void outer_reduce_final_kernel(const float* __restrict__ partial, int Ao, int Bi, int nchunk,
                                          float* __restrict__ out, float* __restrict__ colsum, int accumulate) {
  const int Bi1 = Bi + 1, nout = Ao * Bi1;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < nout; idx += gridDim.x * blockDim.x) {
    float t = 0.f;
    for (int c = 0; c < nchunk; ++c) t += partial[static_cast<size_t>(c) * nout + idx];
    const int o = idx / Bi1, i = idx % Bi1;
    float* dst = i < Bi ? (out != nullptr ? out + o * Bi + i : nullptr) : (colsum != nullptr ? colsum + o : nullptr);
    if (dst != nullptr) *dst = accumulate ? *dst + t : t;
  }
}

// ------------------------------------------------------------------------------------------------
// SOCA tail.  S [B][64][64] (Newton-Schulz square root), v[j] = mean_i S[i][j], s = sigmoid(W2 relu(W1 v + b1) + b2).
// mlp: W1[R][64] b1[R] W2[64][R] b2[64] flat (the layout of dfir_soca).
// ------------------------------------------------------------------------------------------------
constexpr int kSocaMaxR = 16;

__device__ __forceinline__ void soca_mlp_eval(const float* __restrict__ S, const float* __restrict__ mlp, int R,
                                              float* v, float* h, float* sg) {  // shared arrays [64], [R], [64]
  const int t = threadIdx.x;
  if (t < 64) {
    float a = 0.f;
    for (int i = 0; i < 64; ++i) a += S[i * 64 + t];
    v[t] = a * (1.f / 64.f);
  }
  __syncthreads();
  if (t < R) {
    float a = mlp[R * 64 + t];
    for (int j = 0; j < 64; ++j) a = fmaf(mlp[t * 64 + j], v[j], a);
    h[t] = fmaxf(a, 0.f);
  }
  __syncthreads();
  if (t < 64) {
    const float* w2 = mlp + R * 64 + R;
    float a = w2[64 * R + t];
    for (int r = 0; r < R; ++r) a = fmaf(w2[t * R + r], h[r], a);
    sg[t] = 1.f / (1.f + expf(-a));
  }
  __syncthreads();
}

__global__ void __launch_bounds__(64) soca_mlp_fwd_kernel(const float* __restrict__ S, const float* __restrict__ mlp, int R,
                                                          float* __restrict__ svec) {
  __shared__ float v[64], h[kSocaMaxR], sg[64];
  soca_mlp_eval(S + static_cast<size_t>(blockIdx.x) * 4096, mlp, R, v, h, sg);
  svec[blockIdx.x * 64 + threadIdx.x] = sg[threadIdx.x];
}

// one CTA per image: dS[b][i][j] = dv[j] / 64 and the image's parameter-gradient contribution part[b][W1 | b1 | W2 | b2];
// soca_mlp_bwd_sum_kernel adds the contributions in batch order (deterministic)
__global__ void __launch_bounds__(64)
soca_mlp_bwd_kernel(const float* __restrict__ S, const float* __restrict__ dsvec, const float* __restrict__ mlp, int R,
                    float* __restrict__ dS, float* __restrict__ part) {
  __shared__ float v[64], h[kSocaMaxR], sg[64], dpre[64], dh[kSocaMaxR];
  const int t = threadIdx.x, b = blockIdx.x;
  const int npar = 2 * R * 64 + R + 64;
  float* g = part + static_cast<size_t>(b) * npar;
  const float* w2 = mlp + R * 64 + R;
  soca_mlp_eval(S + static_cast<size_t>(b) * 4096, mlp, R, v, h, sg);
  dpre[t] = dsvec[b * 64 + t] * sg[t] * (1.f - sg[t]);
  __syncthreads();
  g[R * 64 + R + 64 * R + t] = dpre[t];                                   // db2
  for (int r = 0; r < R; ++r) g[R * 64 + R + t * R + r] = dpre[t] * h[r];  // dW2[t][r]
  if (t < R) {
    float a = 0.f;
    for (int c = 0; c < 64; ++c) a = fmaf(w2[c * R + t], dpre[c], a);
    dh[t] = h[t] > 0.f ? a : 0.f;
    g[R * 64 + t] = dh[t];                                                 // db1
  }
  __syncthreads();
  float dv = 0.f;
  for (int r = 0; r < R; ++r) {
    g[r * 64 + t] = dh[r] * v[t];                                          // dW1[r][t]
    dv = fmaf(mlp[r * 64 + t], dh[r], dv);
  }
  dv *= (1.f / 64.f);
  float* o = dS + static_cast<size_t>(b) * 4096;
  for (int i = 0; i < 64; ++i) o[i * 64 + t] = dv;
}

__global__ void soca_mlp_bwd_sum_kernel(const float* __restrict__ part, float* __restrict__ dmlp, int npar, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npar) return;
  float t = 0.f;
  for (int b = 0; b < B; ++b) t += part[static_cast<size_t>(b) * npar + i];
  dmlp[i] = t;
}

// ------------------------------------------------------------------------------------------------
// LAM backward.  P[b][i][j] = sum_e dO_i[e] X_j[e];  dA = gamma P;  dE' = A (dA - rowsum(A dA));  dE = -dE' (the
// `max - E` shift cancels: rows of dE' sum to zero);  dX_j = dO_j + gamma sum_i A[i][j] dO_i + sum_i (dE[i][j] + dE[j][i]) X_i
// ------------------------------------------------------------------------------------------------
constexpr int kLamMaxNb = 16;

__global__ void __launch_bounds__(256)
lam_bwd_gram_kernel(const float* __restrict__ stack, long long map_stride, const float* __restrict__ dout,
                    float* __restrict__ partial, int N, int HW, int C, int nchunk) {
  __shared__ float red[8][kLamMaxNb];
  const int chunk = blockIdx.x, b = blockIdx.y, i = blockIdx.z;
  const long long HWC = static_cast<long long>(HW) * C;
  const long long per = (HWC + nchunk - 1) / nchunk;
  const long long e0 = chunk * per, e1 = min(HWC, e0 + per);
  float acc[kLamMaxNb];
#pragma unroll
  for (int j = 0; j < kLamMaxNb; ++j) acc[j] = 0.f;
  const float* base = stack + static_cast<size_t>(b) * HWC;
  const float* dbase = dout + static_cast<size_t>(b) * HWC * N;
  for (long long e = e0 + threadIdx.x; e < e1; e += 256) {
    const long long p = e / C;
    const int c = static_cast<int>(e % C);
    const float d = dbase[p * (static_cast<long long>(N) * C) + i * C + c];
#pragma unroll
    for (int j = 0; j < kLamMaxNb; ++j)
      if (j < N) acc[j] = fmaf(d, base[j * map_stride + e], acc[j]);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < kLamMaxNb; ++j) {
    float v = acc[j];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < N) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    partial[((static_cast<size_t>(b) * nchunk + chunk) * N + i) * N + threadIdx.x] = t;
  }
}

// coef[b] = { M1[N][N] = gamma A, M2[N][N] = dE + dE^T };  dgamma = sum_b sum_ij A_ij P_ij (one CTA, batch in order)
__global__ void lam_bwd_coef_kernel(const float* __restrict__ partial, const float* __restrict__ att, float gamma,
                                    float* __restrict__ coef, float* __restrict__ dgamma, int N, int B, int nchunk) {
  __shared__ float P[kLamMaxNb * kLamMaxNb], A[kLamMaxNb * kLamMaxNb], dE[kLamMaxNb * kLamMaxNb], rowdot[kLamMaxNb];
  __shared__ float gsum;
  if (threadIdx.x == 0) gsum = 0.f;
  for (int b = 0; b < B; ++b) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) {
      float t = 0.f;
      for (int c = 0; c < nchunk; ++c) t += partial[(static_cast<size_t>(b) * nchunk + c) * N * N + idx];
      P[idx] = t;
      A[idx] = att[static_cast<size_t>(b) * N * N + idx];
    }
    __syncthreads();
    if (threadIdx.x < N) {
      float t = 0.f;
      for (int j = 0; j < N; ++j) t = fmaf(A[threadIdx.x * N + j], gamma * P[threadIdx.x * N + j], t);
      rowdot[threadIdx.x] = t;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) {
      const int i = idx / N;
      dE[idx] = -A[idx] * (gamma * P[idx] - rowdot[i]);
    }
    __syncthreads();
    float* cf = coef + static_cast<size_t>(b) * 2 * N * N;
    for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) {
      const int i = idx / N, j = idx % N;
      cf[idx] = gamma * A[idx];
      cf[N * N + idx] = dE[idx] + dE[j * N + i];
    }
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int idx = 0; idx < N * N; ++idx) t = fmaf(A[idx], P[idx], t);
      gsum += t;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) *dgamma = gsum;
}

__global__ void __launch_bounds__(256)
lam_bwd_apply_kernel(const float* __restrict__ stack, long long map_stride, const float* __restrict__ dout,
                     const float* __restrict__ coef, float* __restrict__ dstack, long long dmap_stride, int N, int HW,
                     int C) {
  __shared__ float m1[kLamMaxNb * kLamMaxNb], m2[kLamMaxNb * kLamMaxNb];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < N * N; i += 256) {
    m1[i] = coef[static_cast<size_t>(b) * 2 * N * N + i];
    m2[i] = coef[static_cast<size_t>(b) * 2 * N * N + N * N + i];
  }
  __syncthreads();
  const long long per_img = static_cast<long long>(HW) * C;
  for (long long e = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; e < per_img;
       e += static_cast<long long>(gridDim.x) * 256) {
    const long long p = e / C;
    const int c = static_cast<int>(e % C);
    float xv[kLamMaxNb], dv[kLamMaxNb];
    for (int n = 0; n < N; ++n) {
      xv[n] = stack[n * map_stride + static_cast<size_t>(b) * per_img + e];
      dv[n] = dout[(static_cast<size_t>(b) * HW + p) * (static_cast<size_t>(N) * C) + n * C + c];
    }
    for (int j = 0; j < N; ++j) {
      float t = dv[j];
      for (int i = 0; i < N; ++i) t = fmaf(m1[i * N + j], dv[i], fmaf(m2[j * N + i], xv[i], t));
      dstack[j * dmap_stride + static_cast<size_t>(b) * per_img + e] = t;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// CSAM backward.  u = conv3d(x) + bias, sg = sigmoid(u), out = x (1 + gamma sg):
//   du = dout x gamma sg (1 - sg);  dx = dout (1 + gamma sg) + conv3d^T(du);  dgamma = sum dout x sg;  dbias = sum du;
//   dw[tap] = sum du[e] x[e + off(tap)]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
csam_bwd_du_kernel(const float* __restrict__ x, const float* __restrict__ dout, const float* __restrict__ w27, float bias,
                   float gamma, float* __restrict__ du, float* __restrict__ dx, float* __restrict__ partial, int B, int H,
                   int W, int C) {
  __shared__ float ws[27];
  __shared__ float red[8][29];
  if (threadIdx.x < 27) ws[threadIdx.x] = w27[threadIdx.x];
  __syncthreads();
  float acc[29];
#pragma unroll
  for (int k = 0; k < 29; ++k) acc[k] = 0.f;
  const long long n = static_cast<long long>(B) * H * W * C;
  for (long long e = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; e < n;
       e += static_cast<long long>(gridDim.x) * 256) {
    const int c = static_cast<int>(e % C);
    const long long pix = e / C;
    const int xw = static_cast<int>(pix % W);
    const int y = static_cast<int>((pix / W) % H);
    const int b = static_cast<int>(pix / (static_cast<long long>(W) * H));
    float nb[27];
    float u = bias;
#pragma unroll
    for (int dc = 0; dc < 3; ++dc)
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dxx = 0; dxx < 3; ++dxx) {
          const int cc = c + dc - 1, yy = y + dy - 1, xx = xw + dxx - 1;
          const bool in = cc >= 0 && cc < C && yy >= 0 && yy < H && xx >= 0 && xx < W;
          const float v = in ? x[((static_cast<size_t>(b) * H + yy) * W + xx) * C + cc] : 0.f;
          nb[(dc * 3 + dy) * 3 + dxx] = v;
          u = fmaf(ws[(dc * 3 + dy) * 3 + dxx], v, u);
        }
    const float sg = 1.f / (1.f + expf(-u));
    const float g = dout[e], xv = nb[13];
    const float d = g * xv * gamma * sg * (1.f - sg);
    du[e] = d;
    dx[e] = g * fmaf(gamma, sg, 1.f);
#pragma unroll
    for (int k = 0; k < 27; ++k) acc[k] = fmaf(d, nb[k], acc[k]);
    acc[27] += d;
    acc[28] = fmaf(g * xv, sg, acc[28]);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 29; ++k) {
    float v = acc[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 29) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    partial[static_cast<size_t>(blockIdx.x) * 29 + threadIdx.x] = t;
  }
}

__global__ void csam_bwd_final_kernel(const float* __restrict__ partial, int nblocks, float* __restrict__ dw27,
                                      float* __restrict__ dbias, float* __restrict__ dgamma) {
  if (threadIdx.x < 29) {
    float t = 0.f;
    for (int i = 0; i < nblocks; ++i) t += partial[static_cast<size_t>(i) * 29 + threadIdx.x];
    if (threadIdx.x < 27) dw27[threadIdx.x] = t;
    else if (threadIdx.x == 27) *dbias = t;
    else *dgamma = t;
  }
}

__global__ void __launch_bounds__(256)
csam_bwd_dx_kernel(const float* __restrict__ du, const float* __restrict__ w27, float* __restrict__ dx, int B, int H, int W,
                   int C) {
  __shared__ float ws[27];
  if (threadIdx.x < 27) ws[threadIdx.x] = w27[threadIdx.x];
  __syncthreads();
  const long long n = static_cast<long long>(B) * H * W * C;
  for (long long e = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; e < n;
       e += static_cast<long long>(gridDim.x) * 256) {
    const int c = static_cast<int>(e % C);
    const long long pix = e / C;
    const int xw = static_cast<int>(pix % W);
    const int y = static_cast<int>((pix / W) % H);
    const int b = static_cast<int>(pix / (static_cast<long long>(W) * H));
    float acc = dx[e];
#pragma unroll
    for (int dc = 0; dc < 3; ++dc)
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dxx = 0; dxx < 3; ++dxx) {
          const int cc = c - dc + 1, yy = y - dy + 1, xx = xw - dxx + 1;  // the output position this input fed through tap
          if (cc < 0 || cc >= C || yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
          acc = fmaf(ws[(dc * 3 + dy) * 3 + dxx], du[((static_cast<size_t>(b) * H + yy) * W + xx) * C + cc], acc);
        }
    dx[e] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// Region non-local attention backward (forward kernels nl_project / nl_pool / nl_attend of san_han.cu are re-run by the
// host wrapper to rebuild proj and keys).  Per region: s_qk = theta_q . K_k, a = softmax_k, y_q = sum_k a_qk V_k,
// z_q = W y_q + b + x_q.
//   dy_q = W^T dz_q;  D_q = dy_q . y_q;  ds_qk = a_qk (dy_q . V_k - D_q)
//   dtheta_q = sum_k ds_qk K_k        (nl_bwd_query: one thread per query, two passes over the keys)
//   dK_k = sum_q ds_qk theta_q, dV_k = sum_q a_qk dy_q   (nl_bwd_key: one thread per key, queries streamed)
//   max-pool backward: the gradient of a pooled channel goes to the first maximum of its 2x2 window (torch's rule)
//   dx = dz + Wq^T dproj;  the 1x1-conv weight gradients are outer_reduce calls
// ------------------------------------------------------------------------------------------------
struct RegionB { int y0, y1, x0, x1; };
__device__ __forceinline__ RegionB region_of_b(int r, int H, int W) {
  const int H1 = H / 2, W1 = W / 2;
  RegionB g;
  g.y0 = (r & 1) ? H1 : 0; g.y1 = (r & 1) ? H : H1;
  g.x0 = (r & 2) ? W1 : 0; g.x1 = (r & 2) ? W : W1;
  return g;
}

// qbuf [npix][12]: dy[8], mx, den, D, pad;  ybuf [npix][8];  dproj [npix][24] (theta part written here)
__global__ void __launch_bounds__(128)
nl_bwd_query_kernel(const float* __restrict__ dz, const float* __restrict__ proj, const float* __restrict__ keys,
                    const float* __restrict__ wW, float* __restrict__ qbuf, float* __restrict__ ybuf,
                    float* __restrict__ dproj, int H, int W, int nk_max) {
  __shared__ float ks[256 * 16];
  __shared__ float w_s[64 * 8];
  const int r = blockIdx.y, b = blockIdx.z;
  const RegionB g = region_of_b(r, H, W);
  const int rh = g.y1 - g.y0, rw = g.x1 - g.x0;
  const int nq = rh * rw, nk = (rh / 2) * (rw / 2);
  for (int i = threadIdx.x; i < 64 * 8; i += 128) w_s[i] = wW[i];
  const int q = blockIdx.x * 128 + threadIdx.x;
  const bool active = q < nq;
  const int qy = active ? g.y0 + q / rw : g.y0, qx = active ? g.x0 + q % rw : g.x0;
  const size_t pix = (static_cast<size_t>(b) * H + qy) * W + qx;
  float th[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) th[i] = proj[pix * 24 + i];
  float mx = -3.4e38f, den = 0.f, y[8] = {};
  const float* kbase = keys + (static_cast<size_t>(b) * 4 + r) * nk_max * 16;
  for (int k0 = 0; k0 < nk; k0 += 256) {
    const int nkt = min(256, nk - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < nkt * 16; i += 128) ks[i] = kbase[static_cast<size_t>(k0) * 16 + i];
    __syncthreads();
    for (int k = 0; k < nkt; ++k) {
      const float* kk = ks + k * 16;
      float sc = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) sc = fmaf(th[i], kk[i], sc);
      if (sc > mx) {
        const float f = expf(mx - sc);
        den *= f;
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] *= f;
        mx = sc;
      }
      const float e = expf(sc - mx);
      den += e;
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] = fmaf(e, kk[8 + i], y[i]);
    }
  }
  const float inv = nk > 0 ? 1.f / den : 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) y[i] *= inv;
  // dy = W^T dz (w_s is [64][8]), D = dy . y
  float dy[8] = {};
  if (active) {
    const float* pz = dz + pix * 64;
    for (int c = 0; c < 64; ++c) {
      const float d = pz[c];
#pragma unroll
      for (int i = 0; i < 8; ++i) dy[i] = fmaf(w_s[c * 8 + i], d, dy[i]);
    }
  }
  float D = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) D = fmaf(dy[i], y[i], D);
  float dth[8] = {};
  for (int k0 = 0; k0 < nk; k0 += 256) {
    const int nkt = min(256, nk - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < nkt * 16; i += 128) ks[i] = kbase[static_cast<size_t>(k0) * 16 + i];
    __syncthreads();
    for (int k = 0; k < nkt; ++k) {
      const float* kk = ks + k * 16;
      float sc = 0.f, da = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sc = fmaf(th[i], kk[i], sc);
        da = fmaf(dy[i], kk[8 + i], da);
      }
      const float ds = expf(sc - mx) * inv * (da - D);
#pragma unroll
      for (int i = 0; i < 8; ++i) dth[i] = fmaf(ds, kk[i], dth[i]);
    }
  }
  if (!active) return;
  float* qb = qbuf + pix * 12;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    qb[i] = dy[i];
    ybuf[pix * 8 + i] = y[i];
    dproj[pix * 24 + i] = dth[i];
  }
  qb[8] = mx; qb[9] = inv; qb[10] = D; qb[11] = 0.f;
}

__global__ void __launch_bounds__(128)
nl_bwd_key_kernel(const float* __restrict__ proj, const float* __restrict__ keys, const float* __restrict__ qbuf,
                  float* __restrict__ dproj, int H, int W, int nk_max) {
  __shared__ float qs[128 * 20];  // theta[8], dy[8], mx, inv, D, pad
  const int r = blockIdx.y, b = blockIdx.z;
  const RegionB g = region_of_b(r, H, W);
  const int rh = g.y1 - g.y0, rw = g.x1 - g.x0;
  const int nq = rh * rw, pw = rw / 2, nk = (rh / 2) * pw;
  const int k = blockIdx.x * 128 + threadIdx.x;
  const bool active = k < nk;
  float kv[16];
  const float* kp = keys + ((static_cast<size_t>(b) * 4 + r) * nk_max + (active ? k : 0)) * 16;
#pragma unroll
  for (int i = 0; i < 16; ++i) kv[i] = nk > 0 ? kp[i] : 0.f;
  float dK[8] = {}, dV[8] = {};
  for (int q0 = 0; q0 < nq; q0 += 128) {
    const int nqt = min(128, nq - q0);
    __syncthreads();
    for (int i = threadIdx.x; i < nqt * 20; i += 128) {
      const int qq = q0 + i / 20, f = i % 20;
      const size_t pix = (static_cast<size_t>(b) * H + g.y0 + qq / rw) * W + g.x0 + qq % rw;
      qs[i] = f < 8 ? proj[pix * 24 + f] : (f < 19 ? qbuf[pix * 12 + f - 8] : 0.f);
    }
    __syncthreads();
    if (active)
      for (int qq = 0; qq < nqt; ++qq) {
        const float* s = qs + qq * 20;
        float sc = 0.f, da = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          sc = fmaf(s[i], kv[i], sc);
          da = fmaf(s[8 + i], kv[8 + i], da);
        }
        const float a = expf(sc - s[16]) * s[17];
        const float ds = a * (da - s[18]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          dK[i] = fmaf(ds, s[i], dK[i]);
          dV[i] = fmaf(a, s[8 + i], dV[i]);
        }
      }
  }
  if (!active) return;
  const int ky = k / pw, kx = k % pw;
  size_t pixs[4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
    pixs[j] = (static_cast<size_t>(b) * H + g.y0 + 2 * ky + (j >> 1)) * W + g.x0 + 2 * kx + (j & 1);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    int best = 0;
    float bv = proj[pixs[0] * 24 + 8 + i];
#pragma unroll
    for (int j = 1; j < 4; ++j) {
      const float v = proj[pixs[j] * 24 + 8 + i];
      if (v > bv) { bv = v; best = j; }
    }
    const float gval = i < 8 ? dK[i] : dV[i - 8];
#pragma unroll
    for (int j = 0; j < 4; ++j) dproj[pixs[j] * 24 + 8 + i] = j == best ? gval : 0.f;
  }
}

// dx = dz + Wq^T dproj  (wq [24][64])
__global__ void __launch_bounds__(256)
nl_bwd_dx_kernel(const float* __restrict__ dz, const float* __restrict__ dproj, const float* __restrict__ wq,
                 float* __restrict__ dx, long long npix) {
  __shared__ float ws[24 * 64];
  for (int i = threadIdx.x; i < 24 * 64; i += 256) ws[i] = wq[i];
  __syncthreads();
  const long long n = npix * 64;
  for (long long e = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; e < n;
       e += static_cast<long long>(gridDim.x) * 256) {
    const long long p = e >> 6;
    const int c = static_cast<int>(e & 63);
    float acc = dz[e];
    const float* dp = dproj + p * 24;
#pragma unroll
    for (int o = 0; o < 24; ++o) acc = fmaf(ws[o * 64 + c], dp[o], acc);
    dx[e] = acc;
  }
}

}  // namespace

// ================================================================================================ host wrappers
size_t channel_dot_scratch_floats(int B, int C) { return static_cast<size_t>(B) * kDotChunks * C; }

int channel_dot(const float* a, const float* b, float* out, float* total, int accumulate_total, float* scratch, int B,
                long long HW, int C, cudaStream_t s) {
  if (C % 4 != 0 || C < 4 || C > 1024) return DFIR_ERR_ARG;
  if (B <= 0 || HW <= 0) return DFIR_OK;
  const int lanes = 256 / (C / 4);
  channel_dot_partial_kernel<<<dim3(kDotChunks, B), 256, static_cast<size_t>(lanes) * C * 4, s>>>(a, b, scratch, HW, C);
  channel_dot_final_kernel<<<1, 256, 0, s>>>(scratch, out, total, B * C, C, accumulate_total);
  return ok_or_cuda3();
}

constexpr int kOuterChunks = 128;
size_t outer_reduce_scratch_floats(int Ao, int Bi) { return static_cast<size_t>(kOuterChunks) * Ao * (Bi + 1); }

int outer_reduce(const float* A, int Ao, const float* Bm, int Bi, long long npix, float* out, float* colsum,
                 int accumulate, float* scratch, cudaStream_t s) {
  if (Ao * (Bi + 1) > 256 * kOuterMaxPer || Ao < 1 || Bi < 1) return DFIR_ERR_ARG;
  if (npix <= 0) return DFIR_OK;
  const int nchunk = static_cast<int>(std::min<long long>(kOuterChunks, (npix + kOuterTile - 1) / kOuterTile));
  const size_t smem = static_cast<size_t>(kOuterTile) * (Ao + Bi + 1) * 4;
  outer_reduce_partial_kernel<<<nchunk, 256, smem, s>>>(A, Ao, Bm, Bi, npix, scratch, nchunk);
  outer_reduce_final_kernel<<<(Ao * (Bi + 1) + 255) / 256, 256, 0, s>>>(scratch, Ao, Bi, nchunk, out, colsum, accumulate);
  return ok_or_cuda3();
}

int soca_mlp_forward(const float* S, const float* mlp, int R, float* svec, int B, cudaStream_t s) {
  if (R < 1 || R > kSocaMaxR) return DFIR_ERR_ARG;
  if (B <= 0) return DFIR_OK;
  soca_mlp_fwd_kernel<<<B, 64, 0, s>>>(S, mlp, R, svec);
  return ok_or_cuda3();
}

size_t soca_mlp_bwd_scratch_floats(int B, int R) { return static_cast<size_t>(B) * (2 * R * 64 + R + 64); }

int soca_mlp_backward(const float* S, const float* dsvec, const float* mlp, int R, float* dS, float* dmlp, float* scratch,
                      int B, cudaStream_t s) {
  if (R < 1 || R > kSocaMaxR) return DFIR_ERR_ARG;
  if (B <= 0) return DFIR_OK;
  const int npar = 2 * R * 64 + R + 64;
  soca_mlp_bwd_kernel<<<B, 64, 0, s>>>(S, dsvec, mlp, R, dS, scratch);
  soca_mlp_bwd_sum_kernel<<<(npar + 255) / 256, 256, 0, s>>>(scratch, dmlp, npar, B);
  return ok_or_cuda3();
}

size_t lam_bwd_scratch_floats(int B, int N) {
  return static_cast<size_t>(B) * kDotChunks * N * N + static_cast<size_t>(B) * 2 * N * N;
}

int lam_backward(const float* stack, long long map_stride, const float* att, float gamma, const float* dout, float* dstack,
                 long long dmap_stride, float* dgamma, float* scratch, int N, int B, int HW, int C, cudaStream_t s) {
  if (N < 1 || N > kLamMaxNb) return DFIR_ERR_ARG;
  if (B <= 0 || HW <= 0) return DFIR_OK;
  float* partial = scratch;
  float* coef = scratch + static_cast<size_t>(B) * kDotChunks * N * N;
  lam_bwd_gram_kernel<<<dim3(kDotChunks, B, N), 256, 0, s>>>(stack, map_stride, dout, partial, N, HW, C, kDotChunks);
  lam_bwd_coef_kernel<<<1, 128, 0, s>>>(partial, att, gamma, coef, dgamma, N, B, kDotChunks);
  const long long per_img = static_cast<long long>(HW) * C;
  lam_bwd_apply_kernel<<<dim3(static_cast<unsigned>(std::min<long long>((per_img + 255) / 256, 592)), B), 256, 0, s>>>(
      stack, map_stride, dout, coef, dstack, dmap_stride, N, HW, C);
  return ok_or_cuda3();
}

constexpr int kCsamBlocks = 592;
size_t csam_bwd_scratch_floats(int B, int H, int W, int C) {
  return static_cast<size_t>(B) * H * W * C + static_cast<size_t>(kCsamBlocks) * 29;
}

int csam_backward(const float* x, const float* dout, const float* w27, float bias, float gamma, float* dx, float* dw27,
                  float* dbias, float* dgamma, float* scratch, int B, int H, int W, int C, cudaStream_t s) {
  const long long n = static_cast<long long>(B) * H * W * C;
  if (n <= 0) return DFIR_OK;
  float* du = scratch;
  float* partial = scratch + n;
  const int nblocks = static_cast<int>(std::min<long long>((n + 255) / 256, kCsamBlocks));
  csam_bwd_du_kernel<<<nblocks, 256, 0, s>>>(x, dout, w27, bias, gamma, du, dx, partial, B, H, W, C);
  csam_bwd_final_kernel<<<1, 32, 0, s>>>(partial, nblocks, dw27, dbias, dgamma);
  csam_bwd_dx_kernel<<<nblocks, 256, 0, s>>>(du, w27, dx, B, H, W, C);
  return ok_or_cuda3();
}

size_t nonlocal_bwd_scratch_floats(int B, int H, int W) {
  const size_t npix = static_cast<size_t>(B) * H * W;
  return nonlocal_scratch_floats(B, H, W) + npix * (12 + 8 + 24) + outer_reduce_scratch_floats(24, 64);
}

int nonlocal_backward(const float* x, const float* dz, const float* wq, const float* bq, const float* wW, float* dx,
                      float* dwq, float* dbq, float* dwW, float* dbW, int accumulate, float* scratch, int B, int H, int W,
                      int C, cudaStream_t s) {
  if (C != 64) return DFIR_ERR_ARG;
  if (B == 0 || H < 2 || W < 2) return B == 0 ? DFIR_OK : DFIR_ERR_ARG;
  const long long npix = static_cast<long long>(B) * H * W;
  const int nk_max = ((H - H / 2) / 2) * ((W - W / 2) / 2) + 1;
  float* fwd = scratch;                                   // proj [npix][24], keys [B][4][nk_max][16]
  float* proj = fwd;
  float* keys = fwd + npix * 24;
  float* qbuf = scratch + nonlocal_scratch_floats(B, H, W);
  float* ybuf = qbuf + npix * 12;
  float* dproj = ybuf + npix * 8;
  float* red = dproj + npix * 24;
  DFIR_TRY_RC(nonlocal_project_pool(x, wq, bq, proj, keys, B, H, W, s));
  if (cudaMemsetAsync(dproj, 0, static_cast<size_t>(npix) * 24 * 4, s) != cudaSuccess) return DFIR_ERR_CUDA;
  const int nq_max = (H - H / 2) * (W - W / 2);
  nl_bwd_query_kernel<<<dim3((nq_max + 127) / 128, 4, B), 128, 0, s>>>(dz, proj, keys, wW, qbuf, ybuf, dproj, H, W, nk_max);
  nl_bwd_key_kernel<<<dim3(std::max(1, (nk_max + 127) / 128), 4, B), 128, 0, s>>>(proj, keys, qbuf, dproj, H, W, nk_max);
  nl_bwd_dx_kernel<<<static_cast<unsigned>(std::min<long long>((npix * 64 + 255) / 256, 148 * 8)), 256, 0, s>>>(
      dz, dproj, wq, dx, npix);
  if (cudaGetLastError() != cudaSuccess) return DFIR_ERR_CUDA;
  DFIR_TRY_RC(outer_reduce(dz, 64, ybuf, 8, npix, dwW, dbW, accumulate, red, s));
  return outer_reduce(dproj, 24, x, 64, npix, dwq, dbq, accumulate, red, s);
}

}  // namespace dfir

// ================================================================================================ C ABI
using namespace dfir;

namespace {
inline cudaStream_t SS(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline float* FP(void* p) { return reinterpret_cast<float*>(p); }
}  // namespace

extern "C" {

size_t dfir_channel_dot_scratch_bytes(int B, int C) { return channel_dot_scratch_floats(B, C) * 4; }
int dfir_channel_dot(const float* a, const float* b, float* out, float* total, int accumulate_total, void* scratch,
                     size_t scratch_bytes, int B, long long HW, int C, void* stream) {
  if (a == nullptr || b == nullptr || B < 0) return DFIR_ERR_ARG;
  if (scratch == nullptr || scratch_bytes < dfir_channel_dot_scratch_bytes(B, C)) return DFIR_ERR_WORKSPACE;
  return channel_dot(a, b, out, total, accumulate_total, FP(scratch), B, HW, C, SS(stream));
}

int dfir_soca_mlp(const float* S, const float* mlp, int R, float* svec, int B, void* stream) {
  if (S == nullptr || mlp == nullptr || svec == nullptr) return DFIR_ERR_ARG;
  return soca_mlp_forward(S, mlp, R, svec, B, SS(stream));
}
size_t dfir_soca_mlp_backward_scratch_bytes(int B, int R) { return soca_mlp_bwd_scratch_floats(B, R) * 4; }
int dfir_soca_mlp_backward(const float* S, const float* grad_svec, const float* mlp, int R, float* grad_S, float* grad_mlp,
                           void* scratch, size_t scratch_bytes, int B, void* stream) {
  if (S == nullptr || grad_svec == nullptr || mlp == nullptr || grad_S == nullptr || grad_mlp == nullptr) return DFIR_ERR_ARG;
  if (scratch == nullptr || scratch_bytes < dfir_soca_mlp_backward_scratch_bytes(B, R)) return DFIR_ERR_WORKSPACE;
  return soca_mlp_backward(S, grad_svec, mlp, R, grad_S, grad_mlp, FP(scratch), B, SS(stream));
}

size_t dfir_lam_backward_scratch_bytes(int B, int N) { return lam_bwd_scratch_floats(B, N) * 4; }
int dfir_lam_backward(const float* stack, long long map_stride, const void* fwd_scratch, float gamma, const float* grad_out,
                      float* grad_stack, long long grad_map_stride, float* grad_gamma, void* scratch, size_t scratch_bytes,
                      int N, int B, int HW, int C, void* stream) {
  if (stack == nullptr || fwd_scratch == nullptr || grad_out == nullptr || grad_stack == nullptr || grad_gamma == nullptr)
    return DFIR_ERR_ARG;
  if (scratch == nullptr || scratch_bytes < dfir_lam_backward_scratch_bytes(B, N)) return DFIR_ERR_WORKSPACE;
  const float* att = reinterpret_cast<const float*>(fwd_scratch) + static_cast<size_t>(B) * 32 * N * N;
  return lam_backward(stack, map_stride, att, gamma, grad_out, grad_stack, grad_map_stride, grad_gamma, FP(scratch), N, B,
                      HW, C, SS(stream));
}

size_t dfir_csam_backward_scratch_bytes(int B, int H, int W, int C) { return csam_bwd_scratch_floats(B, H, W, C) * 4; }
int dfir_csam_backward(const float* x, const float* grad_out, const float* w27, float bias, float gamma, float* grad_x,
                       float* grad_w27, float* grad_bias, float* grad_gamma, void* scratch, size_t scratch_bytes, int B,
                       int H, int W, int C, void* stream) {
  if (x == nullptr || grad_out == nullptr || w27 == nullptr || grad_x == nullptr || grad_w27 == nullptr ||
      grad_bias == nullptr || grad_gamma == nullptr)
    return DFIR_ERR_ARG;
  if (scratch == nullptr || scratch_bytes < dfir_csam_backward_scratch_bytes(B, H, W, C)) return DFIR_ERR_WORKSPACE;
  return csam_backward(x, grad_out, w27, bias, gamma, grad_x, grad_w27, grad_bias, grad_gamma, FP(scratch), B, H, W, C,
                       SS(stream));
}

size_t dfir_nonlocal_backward_scratch_bytes(int B, int H, int W) { return nonlocal_bwd_scratch_floats(B, H, W) * 4; }
int dfir_nonlocal_backward(const float* x, const float* grad_out, const float* w_tpg, const float* b_tpg,
                           const float* w_out, float* grad_x, float* grad_w_tpg, float* grad_b_tpg, float* grad_w_out,
                           float* grad_b_out, int accumulate, void* scratch, size_t scratch_bytes, int B, int H, int W,
                           int C, void* stream) {
  if (x == nullptr || grad_out == nullptr || w_tpg == nullptr || b_tpg == nullptr || w_out == nullptr || grad_x == nullptr ||
      grad_w_tpg == nullptr || grad_b_tpg == nullptr || grad_w_out == nullptr || grad_b_out == nullptr)
    return DFIR_ERR_ARG;
  if (scratch == nullptr || scratch_bytes < dfir_nonlocal_backward_scratch_bytes(B, H, W)) return DFIR_ERR_WORKSPACE;
  return nonlocal_backward(x, grad_out, w_tpg, b_tpg, w_out, grad_x, grad_w_tpg, grad_b_tpg, grad_w_out, grad_b_out,
                           accumulate, FP(scratch), B, H, W, C, SS(stream));
}

int dfir_pack_conv3x3_f32_ex(const float* w, float* out, int cout, int cin, int transpose, void* stream) {
  if (w == nullptr || out == nullptr || cout < 1 || cin < 1) return DFIR_ERR_ARG;
  return pack_f32_multi(nullptr, w, out, 1, cout, cin, transpose, SS(stream));
}

}  // extern "C"
