// vk_sisr_loss.cu — the super-resolution negative ELBO, forward value and gradients in one launch sequence
// (sm_100a).  Replaces loss/ELBO_simple.py:82-138 (elbo_sisr) with its helpers :55-80, the kernel synthesis
// utils/util_sisr.py:26-58 (sigma2kernel), the degradation operator utils/util_sisr.py:127-144
// (conv_multi_kernel_tensor: reflect pad, per-sample blur, down-sampling) and the antialiased cubic resize of
// ResizeRight/resize_right.py:29-76, which is a fixed separable linear operator per (size, scale) and arrives
// here as two dense matrices built by the host (virnet_b200/loss/resize_right.py).
//
// Everything is fp32 on NCHW tensors (the loss is <1 % of a training step; accumulations that feed the
// scalar terms are fp64).  The random draws the reference makes inside the loss are inputs, so that the same
// torch generator state gives the same loss.  Launch sequence (vk_elbo_sisr):
//   1 kernel_fwd    kinfo, draws -> 2x2 covariance -> inverse -> softmax Gaussian kernel [N][k*k]
//   2 blur          B = corr(reflect_pad(mu + sqrt(eps2) z), kernel)                      [N,C,H,W]
//   3,4 resize      O = Rh B Rw^T                                                          [N,C,h,w]
//   5 residual      gO = (alpha0-1)/beta_n (O - x) / count,  S_n = sum (x - O)^2
//   6,7 resize^T    gB = Rh^T gO Rw
//   8 blur^T        gpad = full correlation of gB with the flipped kernel                  [N,C,H+k-1,W+k-1]
//   9 mu_grad       d_mu = (mu - hr)/(eps2 count) + reflect-fold(gpad),  sum (mu - hr)^2
//  10 kernel_wgrad  gk[n][i][j] = sum_{c,y,x} gB[n,c,y,x] zz_pad[n,c,y+i,x+j]
//  11 kernel_bwd    softmax / quadratic form / 2x2 inverse / reparameterisation backward -> d_kinfo, d_sigma, terms
//  12 finalize      the 8 scalar terms
#include <algorithm>
#include <cmath>
#include <cstdio>

#include "../../include/virnet_b200.h"
#include "vk_common.cuh"
#include "vk_host.h"

namespace vk {

constexpr int kBlurTile = 32;     // output tile edge of the blur kernels
constexpr int kMaxK = 31;         // largest supported blur kernel edge

__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// ---- 1: sigma2kernel on the re-parameterised covariance (ELBO_simple.py:66-80, util_sisr.py:26-58) ----
// aux[n][0..7] = v1, v2, sqrt(v1 v2), rho (unclamped), A, B, D (inverse covariance entries), det
__global__ void sisr_kernel_fwd_kernel(const float* __restrict__ kinfo, const float* __restrict__ gamma_draw,
                                       const float* __restrict__ rho_draw, float kappa0, float r2, int K, float center,
                                       float* __restrict__ kernel, float* __restrict__ aux) {
  __shared__ float red[32];
  __shared__ float bc[4];
  const int n = blockIdx.x;
  if (threadIdx.x == 0) {
    const float v1 = kappa0 * kinfo[n * 3 + 0] / gamma_draw[n * 2 + 0];
    const float v2 = kappa0 * kinfo[n * 3 + 1] / gamma_draw[n * 2 + 1];
    const float rho = kinfo[n * 3 + 2] + sqrtf(r2) * rho_draw[n];
    const float s12 = sqrtf(v1) * sqrtf(v2);
    const float dir = s12 * fminf(fmaxf(rho, -1.f), 1.f);
    float a = v1, d = v2;
    float det = a * d - dir * dir;
    if (!(det > 0.f)) {               // singular: the reference retries with + 1e-5 I (util_sisr.py:38-40)
      a += 1e-5f, d += 1e-5f;
      det = a * d - dir * dir;
    }
    const float A = d / det, B = -dir / det, D = a / det;
    float* ax = aux + n * 8;
    ax[0] = v1, ax[1] = v2, ax[2] = s12, ax[3] = rho, ax[4] = A, ax[5] = B, ax[6] = D, ax[7] = det;
    bc[0] = A, bc[1] = B, bc[2] = D;
  }
  __syncthreads();
  const float A = bc[0], B = bc[1], D = bc[2];
  const int KK = K * K;
  // softmax over the K*K grid: X = row index, Y = column index (torch.meshgrid 'ij', util_sisr.py:49-50)
  float mx = -INFINITY;
  for (int t = threadIdx.x; t < KK; t += blockDim.x) {
    const float zx = float(t / K) - center, zy = float(t % K) - center;
    mx = fmaxf(mx, -0.5f * (zx * zx * A + 2.f * zx * zy * B + zy * zy * D));
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int i = 1; i < int(blockDim.x >> 5); ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int t = threadIdx.x; t < KK; t += blockDim.x) {
    const float zx = float(t / K) - center, zy = float(t % K) - center;
    const float e = expf(-0.5f * (zx * zx * A + 2.f * zx * zy * B + zy * zy * D) - mx);
    kernel[n * KK + t] = e;
    sum += e;
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
  for (int i = 0; i < int(blockDim.x >> 5); ++i) sum += red[i];
  const float inv = 1.f / sum;
  for (int t = threadIdx.x; t < KK; t += blockDim.x) kernel[n * KK + t] *= inv;
}

// ---- 2 / 8: per-sample KxK cross-correlation of NCHW planes ----
// out[p][y][x] = sum_{i,j} kern[n][i][j] * src(p, y + i - off, x + j - off),  src = in (+ nscale * noise),
// out-of-range reads are reflected (kReflect: the forward blur on the reflect-padded image, off = K/2) or zero
// (the transposed blur: off = K-1, flipped kernel, output grown by K-1).
template <bool kReflect>
__global__ void __launch_bounds__(256)
sisr_blur_kernel(const float* __restrict__ in, const float* __restrict__ noise, float nscale,
                 const float* __restrict__ kern, float* __restrict__ out, int C, int Hin, int Win, int Hout, int Wout,
                 int off, int K, int flip, int flags = 0) {      // flags: 1 = half-sample symmetric padding, 2 = clip to [0, 1]
  extern __shared__ float sm[];
  const int TW = kBlurTile + K - 1, TP = TW + 1;
  float* tile = sm;                      // [TW][TP]
  float* ks = sm + TW * TP;              // [K*K]
  const int p = blockIdx.z, n = p / C;
  const int oy0 = blockIdx.y * kBlurTile, ox0 = blockIdx.x * kBlurTile;
  const float* ip = in + static_cast<long long>(p) * Hin * Win;
  const float* np = noise ? noise + static_cast<long long>(p) * Hin * Win : nullptr;
  const int KK = K * K;
  for (int t = threadIdx.x; t < KK; t += 256) ks[t] = kern[n * KK + (flip ? KK - 1 - t : t)];
  for (int t = threadIdx.x; t < TW * TW; t += 256) {
    const int r = t / TW, c = t % TW;
    int gy = oy0 + r - off, gx = ox0 + c - off;
    float v = 0.f;
    if (kReflect) {
      // rows/cols beyond the last output tile's needs may reflect twice on tiny images: clamp after reflecting
      if (flags & 1) {              // scipy.ndimage 'reflect' (d c b a | a b c d): the edge sample is repeated
        gy = gy < 0 ? -gy - 1 : (gy >= Hin ? 2 * Hin - 1 - gy : gy);
        gx = gx < 0 ? -gx - 1 : (gx >= Win ? 2 * Win - 1 - gx : gx);
        gy = min(max(gy, 0), Hin - 1), gx = min(max(gx, 0), Win - 1);
      } else {
        gy = min(max(reflect_idx(gy, Hin), 0), Hin - 1);
        gx = min(max(reflect_idx(gx, Win), 0), Win - 1);
      }
      v = ip[gy * Win + gx];
      if (np) v = fmaf(nscale, np[gy * Win + gx], v);
    } else if (gy >= 0 && gy < Hin && gx >= 0 && gx < Win) {
      v = ip[gy * Win + gx];
    }
    tile[r * TP + c] = v;
  }
  __syncthreads();
  // each thread owns 4 consecutive columns of one row: a tap row is a sliding window over K + 3 staged values, so
  // a step costs 2 shared loads (one value, one broadcast weight) for 4 FMAs; lanes of a warp touch 32 distinct banks
  const int cg = threadIdx.x & 7, ty = threadIdx.x >> 3;      // 8 column groups x 32 rows
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int i = 0; i < K; ++i) {
    const float* T = tile + (ty + i) * TP + cg * 4;
    const float* kr = ks + i * K;
    float v0 = T[0], v1 = T[1], v2 = T[2], v3;
    int j = 0;
    for (; j + 3 < K; j += 4) {
      float k;
      v3 = T[j + 3], k = kr[j];
      a0 = fmaf(k, v0, a0), a1 = fmaf(k, v1, a1), a2 = fmaf(k, v2, a2), a3 = fmaf(k, v3, a3);
      v0 = T[j + 4], k = kr[j + 1];
      a0 = fmaf(k, v1, a0), a1 = fmaf(k, v2, a1), a2 = fmaf(k, v3, a2), a3 = fmaf(k, v0, a3);
      v1 = T[j + 5], k = kr[j + 2];
      a0 = fmaf(k, v2, a0), a1 = fmaf(k, v3, a1), a2 = fmaf(k, v0, a2), a3 = fmaf(k, v1, a3);
      v2 = T[j + 6], k = kr[j + 3];
      a0 = fmaf(k, v3, a0), a1 = fmaf(k, v0, a1), a2 = fmaf(k, v1, a2), a3 = fmaf(k, v2, a3);
    }
    for (; j < K; ++j) {
      v3 = T[j + 3];
      const float k = kr[j];
      a0 = fmaf(k, v0, a0), a1 = fmaf(k, v1, a1), a2 = fmaf(k, v2, a2), a3 = fmaf(k, v3, a3);
      v0 = v1, v1 = v2, v2 = v3;
    }
  }
  float* op = out + static_cast<long long>(p) * Hout * Wout;
  const int y = oy0 + ty, x = ox0 + cg * 4;
  if (flags & 2) a0 = fminf(fmaxf(a0, 0.f), 1.f), a1 = fminf(fmaxf(a1, 0.f), 1.f), a2 = fminf(fmaxf(a2, 0.f), 1.f),
                 a3 = fminf(fmaxf(a3, 0.f), 1.f);
  if (y < Hout) {
    if (x < Wout) op[y * Wout + x] = a0;
    if (x + 1 < Wout) op[y * Wout + x + 1] = a1;
    if (x + 2 < Wout) op[y * Wout + x + 2] = a2;
    if (x + 3 < Wout) op[y * Wout + x + 3] = a3;
  }
}

// ---- 3,4,6,7: C[p] (M x N) = A (M x Kd) * B (Kd x N), either operand shared by all planes (plane stride 0) ----
// 64 x 64 block tile, 4 x 4 outputs per thread, K step 16; operands are addressed through (row, column) strides so the
// same kernel applies the down-sampling operator and its transpose from either side.
__global__ void __launch_bounds__(256)
sisr_plane_gemm_kernel(const float* __restrict__ A, long long a_plane, int a_rs, int a_cs, const float* __restrict__ B,
                       long long b_plane, int b_rs, int b_cs, float* __restrict__ Cm, long long c_plane, int M, int N,
                       int Kd) {
  __shared__ __align__(16) float as[16][68], bs[16][68];
  const int p = blockIdx.z;
  const float* a = A + a_plane * p;
  const float* b = B + b_plane * p;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < Kd; k0 += 16) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = threadIdx.x + 256 * q;
      {  // A tile: consecutive threads walk k (contiguous for row-major A) unless A is accessed transposed
        const int m = a_cs == 1 ? e >> 4 : e & 63, k = a_cs == 1 ? e & 15 : e >> 6;
        as[k][m] = (m0 + m < M && k0 + k < Kd)
                       ? a[static_cast<long long>(m0 + m) * a_rs + static_cast<long long>(k0 + k) * a_cs] : 0.f;
      }
      {
        const int k = b_cs == 1 ? e >> 6 : e & 15, nn = b_cs == 1 ? e & 63 : e >> 4;
        bs[k][nn] = (k0 + k < Kd && n0 + nn < N)
                        ? b[static_cast<long long>(k0 + k) * b_rs + static_cast<long long>(n0 + nn) * b_cs] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&as[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&bs[k][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* c = Cm + c_plane * p;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col < N) c[static_cast<long long>(row) * N + col] = acc[i][j];
    }
  }
}

// Every reduction of this loss goes through per-block slots that a later step adds in a fixed order: no atomics, so the
// terms and all gradients are bit-identical run to run.
// ---- 5: likelihood residual (ELBO_simple.py:58): O -> gO in place, per-sample sum of squares ----
__global__ void sisr_residual_kernel(float* __restrict__ O, const float* __restrict__ x, const float* __restrict__ sigma_est,
                                     float alpha0, int per_sample, float inv_count, double* __restrict__ ssq_part) {
  __shared__ double red[32];
  const int n = blockIdx.y;
  const float coef = (alpha0 - 1.f) / (sigma_est[n] * alpha0) * inv_count;
  double s = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_sample; i += gridDim.x * blockDim.x) {
    const long long g = static_cast<long long>(n) * per_sample + i;
    const float r = O[g] - x[g];
    s += double(r) * double(r);
    O[g] = coef * r;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < int(blockDim.x >> 5); ++i) t += red[i];
    ssq_part[n * gridDim.x + blockIdx.x] = t;            // one slot per block, summed in block order by step 11
  }
}

// ---- 9: d_mu = Gaussian-KL gradient + reflect-fold of the padded blur gradient; sum (mu - hr)^2 ----
__global__ void sisr_mu_grad_kernel(const float* __restrict__ mu, const float* __restrict__ hr,
                                    const float* __restrict__ gpad, float* __restrict__ d_mu, int H, int W, int pad,
                                    float kl_coef, long long total, double* __restrict__ sq_part) {
  __shared__ double red[32];
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  double s = 0.0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int X = int(i % W), Y = int((i / W) % H);
    const long long p = i / (static_cast<long long>(W) * H);
    const float d = mu[i] - hr[i];
    s += double(d) * double(d);
    // padded positions that read pixel Y under reflect padding: itself, and its mirror images in the borders
    int ys[3], xs[3], ny = 0, nx = 0;
    ys[ny++] = Y + pad;
    if (Y >= 1 && Y <= pad) ys[ny++] = pad - Y;
    if (Y >= H - 1 - pad && Y <= H - 2) ys[ny++] = 2 * (H - 1) - Y + pad;
    xs[nx++] = X + pad;
    if (X >= 1 && X <= pad) xs[nx++] = pad - X;
    if (X >= W - 1 - pad && X <= W - 2) xs[nx++] = 2 * (W - 1) - X + pad;
    const float* gp = gpad + p * Hp * Wp;
    float g = 0.f;
    for (int a = 0; a < ny; ++a)
      for (int b = 0; b < nx; ++b) g += gp[ys[a] * Wp + xs[b]];
    d_mu[i] = fmaf(kl_coef, d, g);
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < int(blockDim.x >> 5); ++i) t += red[i];
    sq_part[blockIdx.x] = t;
  }
}

// ---- 10: gradient w.r.t. the blur kernel, one partial per (plane, 32x32 tile):
//          gk_part[n][slot][i][j] = sum over the tile of gB * zz_pad(y+i, x+j),  slot = (channel, tile row, tile column) ----
__global__ void __launch_bounds__(256)
sisr_kernel_wgrad_kernel(const float* __restrict__ mu, const float* __restrict__ noise, float nscale,
                         const float* __restrict__ gB, float* __restrict__ gk_part, int C, int H, int W, int K) {
  extern __shared__ float sm[];
  const int TW = kBlurTile + K - 1, TP = TW + 1;
  float* tile = sm;                          // [TW][TP] zz with halo
  float* gs = sm + TW * TP;                  // [32][32]
  const int p = blockIdx.z, n = p / C, pad = K / 2;
  const int oy0 = blockIdx.y * kBlurTile, ox0 = blockIdx.x * kBlurTile;
  const float* ip = mu + static_cast<long long>(p) * H * W;
  const float* np = noise + static_cast<long long>(p) * H * W;
  const float* gp = gB + static_cast<long long>(p) * H * W;
  for (int t = threadIdx.x; t < TW * TW; t += 256) {
    const int r = t / TW, c = t % TW;
    const int gy = min(max(reflect_idx(oy0 + r - pad, H), 0), H - 1);
    const int gx = min(max(reflect_idx(ox0 + c - pad, W), 0), W - 1);
    tile[r * TP + c] = fmaf(nscale, np[gy * W + gx], ip[gy * W + gx]);
  }
  for (int t = threadIdx.x; t < kBlurTile * kBlurTile; t += 256) {
    const int y = oy0 + t / kBlurTile, x = ox0 + t % kBlurTile;
    gs[t] = (y < H && x < W) ? gp[y * W + x] : 0.f;
  }
  for (int t = threadIdx.x; t < TW; t += 256) tile[t * TP + TW] = 0.f;
  __syncthreads();
  // a thread accumulates two horizontally adjacent taps (i, 2jj) and (i, 2jj + 1): they read the same staged row
  // shifted by one, so a step is 2 shared loads (one broadcast gradient, one value) for 2 FMAs
  const int KK = K * K, KH = (K + 1) / 2;
  const int tiles_per_plane = gridDim.x * gridDim.y;
  float* gk = gk_part + (static_cast<long long>(p) * tiles_per_plane + blockIdx.y * gridDim.x + blockIdx.x) * KK;   // p = n * C + channel
  for (int w = threadIdx.x; w < K * KH; w += 256) {
    const int i = w / KH, j = (w - i * KH) * 2;
    float acc0 = 0.f, acc1 = 0.f;
    for (int y = 0; y < kBlurTile; ++y) {
      const float* trow = tile + (y + i) * TP + j;
      const float* grow = gs + y * kBlurTile;
      float t0 = trow[0];
#pragma unroll 8
      for (int x = 0; x < kBlurTile; ++x) {
        const float t1 = trow[x + 1], gv = grow[x];          // column TW of the last pair is the (finite) row padding
        acc0 = fmaf(gv, t0, acc0);
        acc1 = fmaf(gv, t1, acc1);
        t0 = t1;
      }
    }
    gk[i * K + j] = acc0;
    if (j + 1 < K) gk[i * K + j + 1] = acc1;
  }
}

// ---- 10b: gk[n][t] = sum of the `slots` partials of sample n in a fixed order: block = 8 warps x 32 consecutive t,
//           warp w adds slots w, w + 8, ... (coalesced rows), the eight warp sums are then added in warp order ----
__global__ void __launch_bounds__(256)
sisr_kernel_wgrad_reduce_kernel(const float* __restrict__ gk_part, float* __restrict__ gk, int slots, int KK) {
  __shared__ float part[8][32];
  const int n = blockIdx.y, w = threadIdx.x >> 5, t = blockIdx.x * 32 + (threadIdx.x & 31);
  const float* src = gk_part + static_cast<long long>(n) * slots * KK;
  float a = 0.f;
  if (t < KK)
    for (int s = w; s < slots; s += 8) a += src[static_cast<long long>(s) * KK + t];
  part[w][threadIdx.x & 31] = a;
  __syncthreads();
  if (w == 0 && t < KK) {
    float v = part[0][threadIdx.x];
#pragma unroll
    for (int j = 1; j < 8; ++j) v += part[j][threadIdx.x];
    gk[n * KK + t] = v;
  }
}

// ---- 11: backward of kernel synthesis + all per-sample scalar terms ----
// per (fp64) [n][5]: lh_n, kl_snet_n, kl_k{0,1,2}_n — summed over n in order by step 12
__global__ void sisr_kernel_bwd_kernel(const float* __restrict__ kernel, const float* __restrict__ gk,
                                       const float* __restrict__ aux, const float* __restrict__ kinfo,
                                       const float* __restrict__ kinfo_gt, const float* __restrict__ gamma_draw,
                                       const float* __restrict__ sigma_est, const float* __restrict__ prior_mean,
                                       const float* __restrict__ prior_logmean, const double* __restrict__ ssq_part,
                                       int ssq_slots, int N,
                                       int K, float center, float kappa0, float r2, float pk0, float pk1, float alpha0,
                                       float digamma_am1, float lr_per_sample, float* __restrict__ d_kinfo,
                                       float* __restrict__ d_sigma, double* __restrict__ per) {
  __shared__ float red[4][32];
  const int n = blockIdx.x, KK = K * K;
  const float* kp = kernel + n * KK;
  const float* gp = gk + n * KK;
  float dot = 0.f;
  for (int t = threadIdx.x; t < KK; t += blockDim.x) dot = fmaf(kp[t], gp[t], dot);
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = dot;
  __syncthreads();
  dot = 0.f;
  for (int i = 0; i < int(blockDim.x >> 5); ++i) dot += red[0][i];
  __syncthreads();
  // softmax backward, then the quadratic form q = -0.5 (zx^2 A + 2 zx zy B + zy^2 D)
  float gA = 0.f, gO = 0.f, gD = 0.f;
  for (int t = threadIdx.x; t < KK; t += blockDim.x) {
    const float gq = kp[t] * (gp[t] - dot);
    const float zx = float(t / K) - center, zy = float(t % K) - center;
    gA = fmaf(-0.5f * zx * zx, gq, gA);
    gO = fmaf(-0.5f * zx * zy, gq, gO);       // each of the two off-diagonal entries of the inverse
    gD = fmaf(-0.5f * zy * zy, gq, gD);
  }
  for (int o = 16; o > 0; o >>= 1) {
    gA += __shfl_xor_sync(0xffffffffu, gA, o);
    gO += __shfl_xor_sync(0xffffffffu, gO, o);
    gD += __shfl_xor_sync(0xffffffffu, gD, o);
  }
  if ((threadIdx.x & 31) == 0) red[1][threadIdx.x >> 5] = gA, red[2][threadIdx.x >> 5] = gO, red[3][threadIdx.x >> 5] = gD;
  __syncthreads();
  if (threadIdx.x != 0) return;
  gA = gO = gD = 0.f;
  for (int i = 0; i < int(blockDim.x >> 5); ++i) gA += red[1][i], gO += red[2][i], gD += red[3][i];
  const float* ax = aux + n * 8;
  const float s12 = ax[2], rho = ax[3], A = ax[4], B = ax[5], D = ax[6];
  // inverse backward: g_cov = -S^T G S^T with S = [[A,B],[B,D]] (symmetric), G = [[gA,gO],[gO,gD]]
  const float m00 = gA * A + gO * B, m01 = gA * B + gO * D, m10 = gO * A + gD * B, m11 = gO * B + gD * D;   // G S
  const float c00 = -(A * m00 + B * m10), c01 = -(A * m01 + B * m11), c10 = -(B * m00 + D * m10),
              c11 = -(B * m01 + D * m11);                                                                  // -S (G S)
  const float g_v1 = c00, g_v2 = c11, g_dir = c01 + c10;
  const float g_rho = (rho >= -1.f && rho <= 1.f) ? g_dir * s12 : 0.f;        // v1, v2 detached in `direction`
  const float k0 = kinfo[n * 3 + 0], k1 = kinfo[n * 3 + 1], k2 = kinfo[n * 3 + 2];
  const float t0 = kinfo_gt[n * 3 + 0], t1 = kinfo_gt[n * 3 + 1], t2 = kinfo_gt[n * 3 + 2];
  const float invN = 1.f / float(N), kscale = pk1 / 3.f * invN;
  // kl_knet (ELBO_simple.py:118-121): inverse-Gamma KLs on the two variances, Gaussian KL on rho
  d_kinfo[n * 3 + 0] = g_v1 * kappa0 / gamma_draw[n * 2 + 0] + kscale * (kappa0 - 1.f) * (1.f / k0 - t0 / (k0 * k0));
  d_kinfo[n * 3 + 1] = g_v2 * kappa0 / gamma_draw[n * 2 + 1] + kscale * (kappa0 - 1.f) * (1.f / k1 - t1 / (k1 * k1));
  d_kinfo[n * 3 + 2] = g_rho + kscale * pk0 * (k2 - t2) / r2;
  per[n * 5 + 2] = double((kappa0 - 1.f) * ((t0 / k0 - 1.f) + (logf(kappa0 * k0) - logf(kappa0 * t0))));
  per[n * 5 + 3] = double((kappa0 - 1.f) * ((t1 / k1 - 1.f) + (logf(kappa0 * k1) - logf(kappa0 * t1))));
  per[n * 5 + 4] = double((k2 - t2) * (k2 - t2));
  // noise variance: likelihood + inverse-Gamma KL (ELBO_simple.py:113-115, :58)
  const float beta = sigma_est[n] * alpha0, am1 = alpha0 - 1.f;
  const float b0m = prior_mean[n] * alpha0, lb0m = logf(alpha0) + prior_logmean[n];
  double ssq = 0.0;
  for (int i = 0; i < ssq_slots; ++i) ssq += ssq_part[n * ssq_slots + i];
  const float msq = float(ssq / double(lr_per_sample));
  per[n * 5 + 0] = double(0.9189385332046727f + 0.5f * (logf(beta) - digamma_am1) + 0.5f * am1 / beta * msq);
  per[n * 5 + 1] = double(am1 * ((b0m / beta - 1.f) + (logf(beta) - lb0m)));
  const float dbeta = 0.5f / beta - 0.5f * am1 * msq / (beta * beta) + am1 * (1.f / beta - b0m / (beta * beta));
  d_sigma[n] = alpha0 * dbeta * invN;
}

// ---- 12: terms = [loss, lh, kl_rnet, kl_snet, kl_knet, kl_knet0, kl_knet1, kl_knet2] ----
// one block of 256 threads: thread t adds the (mu - hr)^2 partials t, t + 256, ... in order, then thread 0 adds the 256
// sums and the per-sample scalars in index order
__global__ void sisr_finalize_kernel(const double* __restrict__ per, const double* __restrict__ sq_part, int sq_slots,
                                     int N, double hr_count, float eps2, float r2, float pk0, float pk1,
                                     float* __restrict__ terms) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < sq_slots; i += 256) s += sq_part[i];
  red[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x != 0) return;
  double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int i = 0; i < 256; ++i) acc[1] += red[i];
  for (int n = 0; n < N; ++n) {
    acc[0] += per[n * 5 + 0], acc[2] += per[n * 5 + 1];
    acc[3] += per[n * 5 + 2], acc[4] += per[n * 5 + 3], acc[5] += per[n * 5 + 4];
  }
  const double lh = acc[0] / N, kl_r = 0.5 * acc[1] / (double(eps2) * hr_count), kl_s = acc[2] / N;
  const double k0 = acc[3] / N, k1 = acc[4] / N, k2 = 0.5 * acc[5] / (double(r2) * N) * pk0;
  const double kl_k = (k0 + k1 + k2) / 3.0 * pk1;
  terms[0] = float(lh + kl_r + kl_s + kl_k);
  terms[1] = float(lh), terms[2] = float(kl_r), terms[3] = float(kl_s), terms[4] = float(kl_k);
  terms[5] = float(k0), terms[6] = float(k1), terms[7] = float(k2);
}

// im_lr = clip(im_blur + noise * std[n], 0, 1) (datasets/SISRDatasets.py:100-104)
__global__ void sisr_add_noise_kernel(const float* __restrict__ blur, const float* __restrict__ noise,
                                      const float* __restrict__ std, float* __restrict__ out, int per_sample, long long total) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = __fadd_rn(blur[i], __fmul_rn(noise[i], std[i / per_sample]));
    out[i] = fminf(fmaxf(v, 0.f), 1.f);
  }
}

}  // namespace vk

using namespace vk;

namespace {
constexpr int kResidualBlocks = 64;          // blocks per sample of step 5 (slots of the per-sample sum of squares)
constexpr int kMuGradBlocks = 148 * 16;      // blocks of step 9 (slots of sum (mu - hr)^2)
struct SisrWs {
  size_t b, t, o, gpad, gk, gk_part, aux, acc, total;
};
int blur_tiles(long long v) { return int((v + kBlurTile - 1) / kBlurTile); }
SisrWs sisr_ws_layout(long long n, long long c, long long H, long long W, long long h, long long w, long long k) {
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  SisrWs L{};
  size_t o = 0;
  L.b = o, o += al(size_t(n * c * H * W) * 4);                       // blur B, later gB
  L.t = o, o += al(size_t(n * c * H * w) * 4);                       // B Rw^T, later Rh^T gO
  L.o = o, o += al(size_t(n * c * h * w) * 4);                       // O, then gO in place
  L.gpad = o, o += al(size_t(n * c * (H + k - 1) * (W + k - 1)) * 4);
  L.gk = o, o += al(size_t(n * k * k) * 4);
  L.gk_part = o, o += al(size_t(n * c * blur_tiles(H) * blur_tiles(W) * k * k) * 4);   // step 10's per-tile partials
  L.aux = o, o += al(size_t(n * 8) * 4);
  // fp64: per-sample scalars [n][5], step 5's slots [n][kResidualBlocks], step 9's slots [kMuGradBlocks]
  L.acc = o, o += al(size_t(n * 5 + n * kResidualBlocks + kMuGradBlocks) * 8);
  L.total = o;
  return L;
}
}  // namespace

extern "C" int64_t vk_elbo_sisr_ws_bytes(int32_t n, int32_t c, int32_t H, int32_t W, int32_t h, int32_t w, int32_t k) {
  if (n <= 0 || c <= 0 || H <= 0 || W <= 0 || h <= 0 || w <= 0 || k <= 0) return -1;
  return int64_t(sisr_ws_layout(n, c, H, W, h, w, k).total);
}

extern "C" uint32_t vk_sizeof_elbo_sisr_args(void) { return uint32_t(sizeof(vk_elbo_sisr_args)); }

extern "C" int vk_elbo_sisr(const vk_elbo_sisr_args* a, void* stream_) {
  if (!a) return VK_E_BADARG;
  const int N = a->n, C = a->c, H = a->H, W = a->W, h = a->h, w = a->w, K = a->k_size;
  if (!a->mu || !a->im_hr || !a->im_lr || !a->sigma_est || !a->kinfo_est || !a->kinfo_gt || !a->prior_mean ||
      !a->prior_logmean || !a->gamma_draw || !a->rho_draw || !a->z_draw || !a->rh || !a->rw || !a->d_mu ||
      !a->d_sigma || !a->d_kinfo || !a->kernel || !a->terms || !a->ws)
    return VK_E_BADARG;
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || h <= 0 || w <= 0 || K < 1 || K > kMaxK || (K & 1) == 0) return VK_E_BADARG;
  if (K / 2 >= H || K / 2 >= W) return VK_E_BADARG;                  // reflect padding needs pad < size
  const SisrWs L = sisr_ws_layout(N, C, H, W, h, w, K);
  if (a->ws_bytes < int64_t(L.total)) return VK_E_BADARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  char* base = reinterpret_cast<char*>(a->ws);
  float* Bf = reinterpret_cast<float*>(base + L.b);
  float* Tf = reinterpret_cast<float*>(base + L.t);
  float* Of = reinterpret_cast<float*>(base + L.o);
  float* gpad = reinterpret_cast<float*>(base + L.gpad);
  float* gk = reinterpret_cast<float*>(base + L.gk);
  float* gk_part = reinterpret_cast<float*>(base + L.gk_part);
  float* aux = reinterpret_cast<float*>(base + L.aux);
  double* per = reinterpret_cast<double*>(base + L.acc);
  double* ssq_part = per + size_t(N) * 5;
  double* sq_part = ssq_part + size_t(N) * kResidualBlocks;
  const int P = N * C, KK = K * K, pad = K / 2;
  const float nscale = sqrtf(a->eps2);
  int launches = 0;

  sisr_kernel_fwd_kernel<<<N, 256, 0, st>>>(a->kinfo_est, a->gamma_draw, a->rho_draw, a->kappa0, a->r2, K, a->center,
                                            a->kernel, aux);
  ++launches;
  const int TW = kBlurTile + K - 1;
  const size_t blur_smem = (size_t(TW) * (TW + 1) + KK) * 4;
  const size_t wg_smem = (size_t(TW) * (TW + 1) + kBlurTile * kBlurTile) * 4;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(sisr_blur_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(sisr_blur_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(sisr_kernel_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    attr_done = true;
  }
  auto tiles = [](int v) { return (v + kBlurTile - 1) / kBlurTile; };
  // 2: B = blur(mu + sqrt(eps2) z)
  sisr_blur_kernel<true><<<dim3(tiles(W), tiles(H), P), 256, blur_smem, st>>>(a->mu, a->z_draw, nscale, a->kernel, Bf, C, H,
                                                                             W, H, W, pad, K, 0);
  ++launches;
  // 3: T[p] (H x w) = B[p] (H x W) * Rw^T     4: O[p] (h x w) = Rh (h x H) * T[p]
  auto gemm = [&](const float* A_, long long ap, int ars, int acs, const float* B_, long long bp, int brs, int bcs,
                  float* C_, long long cpl, int M, int Nn, int Kd) {
    sisr_plane_gemm_kernel<<<dim3((Nn + 63) / 64, (M + 63) / 64, P), 256, 0, st>>>(A_, ap, ars, acs, B_, bp, brs, bcs, C_,
                                                                                 cpl, M, Nn, Kd);
    ++launches;
  };
  gemm(Bf, (long long)H * W, W, 1, a->rw, 0, 1, W, Tf, (long long)H * w, H, w, W);
  gemm(a->rh, 0, H, 1, Tf, (long long)H * w, w, 1, Of, (long long)h * w, h, w, H);
  // 5: residual / likelihood gradient w.r.t. the degraded image
  const int per_sample = C * h * w;
  const int res_blocks = std::min((per_sample + 255) / 256, kResidualBlocks);
  sisr_residual_kernel<<<dim3(res_blocks, N), 256, 0, st>>>(Of, a->im_lr, a->sigma_est, a->alpha0, per_sample,
                                                            1.f / (float(N) * float(per_sample)), ssq_part);
  ++launches;
  // 6: T[p] (H x w) = Rh^T (H x h) * gO[p]    7: gB[p] (H x W) = T[p] (H x w) * Rw (w x W)
  gemm(a->rh, 0, 1, H, Of, (long long)h * w, w, 1, Tf, (long long)H * w, H, w, h);
  gemm(Tf, (long long)H * w, w, 1, a->rw, 0, W, 1, Bf, (long long)H * W, H, W, w);
  // 8: transposed blur onto the padded grid
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  sisr_blur_kernel<false><<<dim3(tiles(Wp), tiles(Hp), P), 256, blur_smem, st>>>(Bf, nullptr, 0.f, a->kernel, gpad, C, H, W,
                                                                                Hp, Wp, K - 1, K, 1);
  ++launches;
  // 9: d_mu
  const long long total = (long long)P * H * W;
  const double hr_count = double(total);
  const int mu_blocks = int(std::min<long long>((total + 255) / 256, kMuGradBlocks));
  sisr_mu_grad_kernel<<<mu_blocks, 256, 0, st>>>(a->mu, a->im_hr, gpad, a->d_mu, H, W, pad,
                                                 float(1.0 / (double(a->eps2) * hr_count)), total, sq_part);
  ++launches;
  // 10: gradient w.r.t. the blur kernel
  sisr_kernel_wgrad_kernel<<<dim3(tiles(W), tiles(H), P), 256, wg_smem, st>>>(a->mu, a->z_draw, nscale, Bf, gk_part, C, H, W,
                                                                             K);
  sisr_kernel_wgrad_reduce_kernel<<<dim3((KK + 31) / 32, N), 256, 0, st>>>(gk_part, gk, C * tiles(W) * tiles(H), KK);
  launches += 2;
  // 11, 12
  sisr_kernel_bwd_kernel<<<N, 256, 0, st>>>(a->kernel, gk, aux, a->kinfo_est, a->kinfo_gt, a->gamma_draw, a->sigma_est,
                                            a->prior_mean, a->prior_logmean, ssq_part, res_blocks, N, K, a->center,
                                            a->kappa0, a->r2, a->pk0, a->pk1, a->alpha0, a->digamma_am1, float(per_sample),
                                            a->d_kinfo, a->d_sigma, per);
  sisr_finalize_kernel<<<1, 256, 0, st>>>(per, sq_part, mu_blocks, N, hr_count, a->eps2, a->r2, a->pk0, a->pk1, a->terms);
  launches += 2;
  g_launch_count.fetch_add(launches, std::memory_order_relaxed);
  return int(cudaGetLastError());
}

// ---- device-side synthesis of SISR training pairs (datasets/SISRDatasets.py:86-104 after the crop / augmentation) ----
extern "C" int64_t vk_sisr_degrade_ws_bytes(int32_t n, int32_t c, int32_t H, int32_t W, int32_t w) {
  if (n <= 0 || c <= 0 || H <= 0 || W <= 0 || w <= 0) return -1;
  return int64_t(n) * c * H * (int64_t(W) + w) * 4 + 512;
}

extern "C" int vk_sisr_degrade(const float* im_hr, const float* kernels, int32_t k_size, const float* rh, const float* rw,
                               const float* noise, const float* std, float* im_blur, float* im_lr, void* ws,
                               int64_t ws_bytes, int32_t n, int32_t c, int32_t H, int32_t W, int32_t h, int32_t w,
                               void* stream_) {
  if (!im_hr || !kernels || !rh || !rw || !noise || !std || !im_blur || !im_lr || !ws) return VK_E_BADARG;
  if (n <= 0 || c <= 0 || H <= 0 || W <= 0 || h <= 0 || w <= 0 || k_size < 1 || k_size > kMaxK || (k_size & 1) == 0)
    return VK_E_BADARG;
  if (k_size / 2 >= H || k_size / 2 >= W || ws_bytes < vk_sisr_degrade_ws_bytes(n, c, H, W, w)) return VK_E_BADARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  float* Bf = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) / 256 * 256);
  float* Tf = Bf + static_cast<long long>(n) * c * H * W;
  const int P = n * c, K = k_size, TW = kBlurTile + K - 1;
  const size_t blur_smem = (size_t(TW) * (TW + 1) + size_t(K) * K) * 4;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(sisr_blur_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    attr_done = true;
  }
  auto tiles = [](int v) { return (v + kBlurTile - 1) / kBlurTile; };
  // scipy.ndimage.convolve(mode='reflect') = correlation with the flipped kernel on symmetric padding; then clip
  sisr_blur_kernel<true><<<dim3(tiles(W), tiles(H), P), 256, blur_smem, st>>>(im_hr, nullptr, 0.f, kernels, Bf, c, H, W, H, W,
                                                                             K / 2, K, 1, 3);
  sisr_plane_gemm_kernel<<<dim3((w + 63) / 64, (H + 63) / 64, P), 256, 0, st>>>(Bf, (long long)H * W, W, 1, rw, 0, 1, W, Tf,
                                                                               (long long)H * w, H, w, W);
  sisr_plane_gemm_kernel<<<dim3((w + 63) / 64, (h + 63) / 64, P), 256, 0, st>>>(rh, 0, H, 1, Tf, (long long)H * w, w, 1,
                                                                               im_blur, (long long)h * w, h, w, H);
  const long long total = static_cast<long long>(P) * h * w;
  sisr_add_noise_kernel<<<int(std::min<long long>((total + 255) / 256, 148 * 8)), 256, 0, st>>>(im_blur, noise, std, im_lr,
                                                                                             c * h * w, total);
  g_launch_count.fetch_add(4, std::memory_order_relaxed);
  return int(cudaGetLastError());
}
