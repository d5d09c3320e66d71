// vk_denoise_real.cu — the two data-side operators of the real-noise denoising trainer (sm_100a):
//   vk_noise_estimate: utils/util_denoising.py:54-63 (noise_estimate_fun) — the inverse-Gamma prior's variance map,
//                      a k x k Gaussian-window mean of (noisy - gt)^2 with reflect padding, clamped at 1e-10;
//   vk_mixup:          datasets/data_tools.py:12-30 (MixUp_AUG.aug) — convex combination of every sample with a
//                      permuted partner, applied to the clean and the noisy batch with the same coefficients.
// Both are HBM-bound, NCHW fp32 in and out (they sit in front of vk_pack_input / vk_elbo_denoise).
#include <algorithm>

#include "../../include/virnet_b200.h"
#include "vk_common.cuh"
#include "vk_host.h"

namespace vk {

constexpr int kNeTile = 32;

// out[p][y][x] = max(floor, sum_ij win[i][j] * err2(p, refl(y+i-r), refl(x+j-r))),  err2 = (noisy - gt)^2
__global__ void __launch_bounds__(256)
noise_estimate_kernel(const float* __restrict__ noisy, const float* __restrict__ gt, const float* __restrict__ win,
                      float* __restrict__ out, int H, int W, int K, float floor_) {
  extern __shared__ float sm[];
  const int TW = kNeTile + K - 1, TP = TW + 1, r = K / 2;
  float* tile = sm;                  // [TW][TP] squared error with halo
  float* ws = sm + TW * TP;          // [K*K]
  const long long pbase = static_cast<long long>(blockIdx.z) * H * W;
  const int oy0 = blockIdx.y * kNeTile, ox0 = blockIdx.x * kNeTile;
  for (int t = threadIdx.x; t < K * K; t += 256) ws[t] = win[t];
  for (int t = threadIdx.x; t < TW * TW; t += 256) {
    int gy = oy0 + t / TW - r, gx = ox0 + t % TW - r;
    gy = gy < 0 ? -gy : gy, gx = gx < 0 ? -gx : gx;
    gy = gy >= H ? 2 * (H - 1) - gy : gy, gx = gx >= W ? 2 * (W - 1) - gx : gx;
    gy = min(max(gy, 0), H - 1), gx = min(max(gx, 0), W - 1);       // only reached by halo cells no output uses
    const float d = noisy[pbase + gy * W + gx] - gt[pbase + gy * W + gx];
    tile[(t / TW) * TP + t % TW] = d * d;
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = 0; i < K; ++i)
    for (int j = 0; j < K; ++j) {
      const float kv = ws[i * K + j];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = fmaf(kv, tile[(ty + 8 * q + i) * TP + tx + j], acc[q]);
    }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int y = oy0 + ty + 8 * q, x = ox0 + tx;
    if (y < H && x < W) out[pbase + y * W + x] = fmaxf(acc[q], floor_);
  }
}

// out_x[n] = lam[n] * x[n] + (1 - lam[n]) * x[perm[n]] for x in {a, b}; float4 over each sample
__global__ void mixup_kernel(const float4* __restrict__ a, const float4* __restrict__ b, const long long* __restrict__ perm,
                             const float* __restrict__ lam, float4* __restrict__ out_a, float4* __restrict__ out_b,
                             long long per_sample4) {
  const int n = blockIdx.y;
  const float l = lam[n], m = 1.f - l;
  const long long src = perm[n] * per_sample4, dst = static_cast<long long>(n) * per_sample4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < per_sample4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 a0 = a[dst + i], a1 = a[src + i], b0 = b[dst + i], b1 = b[src + i];
    out_a[dst + i] = make_float4(l * a0.x + m * a1.x, l * a0.y + m * a1.y, l * a0.z + m * a1.z, l * a0.w + m * a1.w);
    out_b[dst + i] = make_float4(l * b0.x + m * b1.x, l * b0.y + m * b1.y, l * b0.z + m * b1.z, l * b0.w + m * b1.w);
  }
}

}  // namespace vk

using namespace vk;

extern "C" int vk_noise_estimate(const float* noisy, const float* gt, const float* window, int32_t k_size, float* out,
                                 int32_t planes, int32_t h, int32_t w, float floor_, void* stream) {
  if (!noisy || !gt || !window || !out || planes <= 0 || h <= 0 || w <= 0) return VK_E_BADARG;
  if (k_size < 1 || k_size > 31 || (k_size & 1) == 0 || k_size / 2 >= h || k_size / 2 >= w) return VK_E_BADARG;
  const int TW = kNeTile + k_size - 1;
  const size_t smem = (size_t(TW) * (TW + 1) + size_t(k_size) * k_size) * sizeof(float);
  const dim3 grid((w + kNeTile - 1) / kNeTile, (h + kNeTile - 1) / kNeTile, planes);
  noise_estimate_kernel<<<grid, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(noisy, gt, window, out, h, w, k_size,
                                                                                   floor_);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return int(cudaGetLastError());
}

extern "C" int vk_mixup(const float* a, const float* b, const int64_t* perm, const float* lam, float* out_a, float* out_b,
                        int32_t n, int64_t per_sample, void* stream) {
  if (!a || !b || !perm || !lam || !out_a || !out_b || n <= 0 || per_sample <= 0 || per_sample % 4 != 0) return VK_E_BADARG;
  if (out_a == a || out_b == b) return VK_E_BADARG;          // partners are read after their own slot is written
  const long long p4 = per_sample / 4;
  const dim3 grid(unsigned(std::min<long long>((p4 + 255) / 256, 148 * 8 / std::max(1, std::min(n, 8)) + 1)), unsigned(n));
  mixup_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b), reinterpret_cast<const long long*>(perm), lam,
      reinterpret_cast<float4*>(out_a), reinterpret_cast<float4*>(out_b), p4);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return int(cudaGetLastError());
}
