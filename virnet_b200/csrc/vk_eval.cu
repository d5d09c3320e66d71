// vk_eval.cu — evaluation-side kernels of the VIRNet callers, sm_100a, all HBM-bound streaming:
//
//   vk_aug8 / vk_aug8_merge : the 8-fold flip / rotate self-ensemble of the SIDD / DND scripts
//                             (scripts/denoising_virnet_real_sidd.py:120-136, dnd_submission_py/pytorch_wrapper.py:17-32,
//                              utils/util_image.py:391-466 data_aug_np / inverse_data_aug_np) as ONE gather kernel in
//                             front of one batched forward and ONE averaging kernel behind it;
//   vk_to_u8                : skimage.img_as_ubyte(clamp(x, 0, 1)) on NCHW fp32 -> HWC uint8 (fp32 x * 255, rint, as
//                             skimage does for float32 input; scripts/*.py `img_as_ubyte(... .clamp(0, 1) ...)`);
//   vk_psnr_u8              : utils/util_image.py:68-89 calculate_psnr on uint8 images — exact integer sum of squared
//                             differences (optionally on the Y channel of rgb2ycbcr :129-153, border cropped);
//   vk_ssim_u8              : utils/util_image.py:16-66 calculate_ssim (11x11 Gaussian window, sigma 1.5, "valid" region,
//                             fp64 like the reference's cv2.filter2D on float64 images).
#include <algorithm>
#include <cstdio>

#include "../../include/virnet_b200.h"
#include "vk_common.cuh"
#include "vk_host.h"

namespace vk {
namespace {

// source pixel (sy, sx) in the H x W input that lands at (y, x) of augmentation `mode`
// (mode 2, 3, 6, 7 outputs are W x H): np.rot90 / np.flipud algebra of data_aug_np
__device__ __forceinline__ void aug_src(int mode, int y, int x, int H, int W, int& sy, int& sx) {
  switch (mode) {
    case 0: sy = y, sx = x; break;
    case 1: sy = H - 1 - y, sx = x; break;
    case 2: sy = x, sx = W - 1 - y; break;
    case 3: sy = x, sx = y; break;
    case 4: sy = H - 1 - y, sx = W - 1 - x; break;
    case 5: sy = y, sx = W - 1 - x; break;
    case 6: sy = H - 1 - x, sx = y; break;
    default: sy = H - 1 - x, sx = W - 1 - y; break;
  }
}

// out_a [4][P][H][W] = modes {0, 1, 4, 5}; out_b [4][P][W][H] = modes {2, 3, 6, 7}; P = N * C planes
__global__ void aug8_kernel(const float* __restrict__ in, float* __restrict__ out_a, float* __restrict__ out_b,
                            long long planes, int H, int W) {
  const long long hw = static_cast<long long>(H) * W;
  const long long total = planes * hw * 8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long pix = i % hw;
    const long long pl = (i / hw) % planes;
    const int slot = int(i / (hw * planes));                 // 0..3 -> out_a, 4..7 -> out_b
    const bool rot = slot >= 4;
    const int k = slot & 3;
    const int mode = rot ? (k == 0 ? 2 : k == 1 ? 3 : k == 2 ? 6 : 7) : (k == 0 ? 0 : k == 1 ? 1 : k == 2 ? 4 : 5);
    const int ow = rot ? H : W;
    const int y = int(pix / ow), x = int(pix % ow);
    int sy, sx;
    aug_src(mode, y, x, H, W, sy, sx);
    const float v = __ldg(in + pl * hw + static_cast<long long>(sy) * W + sx);
    (rot ? out_b : out_a)[(static_cast<long long>(k) * planes + pl) * hw + pix] = v;
  }
}

// out[p][sy][sx] = mean over the 8 modes of the network output at the augmented position of (sy, sx):
// the inverse of a permutation gather is the same gather read the other way round
__global__ void aug8_merge_kernel(const float* __restrict__ in_a, const float* __restrict__ in_b, float* __restrict__ out,
                                  long long planes, int H, int W, int clip01) {
  const long long hw = static_cast<long long>(H) * W;
  const long long total = planes * hw;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long pl = i / hw;
    const int sy = int((i % hw) / W), sx = int(i % W);
    // positions (y, x) in each augmented image whose source is (sy, sx)
    float acc = 0.f;
    // modes 0, 1, 4, 5 (H x W), in the reference's accumulation order 0..7
    const float m0 = __ldg(in_a + (0 * planes + pl) * hw + static_cast<long long>(sy) * W + sx);
    const float m1 = __ldg(in_a + (1 * planes + pl) * hw + static_cast<long long>(H - 1 - sy) * W + sx);
    const float m4 = __ldg(in_a + (2 * planes + pl) * hw + static_cast<long long>(H - 1 - sy) * W + (W - 1 - sx));
    const float m5 = __ldg(in_a + (3 * planes + pl) * hw + static_cast<long long>(sy) * W + (W - 1 - sx));
    // modes 2, 3, 6, 7 (W x H): invert (sy, sx) = f(y, x)
    const float m2 = __ldg(in_b + (0 * planes + pl) * hw + static_cast<long long>(W - 1 - sx) * H + sy);
    const float m3 = __ldg(in_b + (1 * planes + pl) * hw + static_cast<long long>(sx) * H + sy);
    const float m6 = __ldg(in_b + (2 * planes + pl) * hw + static_cast<long long>(sx) * H + (H - 1 - sy));
    const float m7 = __ldg(in_b + (3 * planes + pl) * hw + static_cast<long long>(W - 1 - sx) * H + (H - 1 - sy));
    acc = m0;
    acc += m1, acc += m2, acc += m3, acc += m4, acc += m5, acc += m6, acc += m7;   // im_denoise += ... ; /= 8
    acc *= 0.125f;
    if (clip01) acc = fminf(fmaxf(acc, 0.f), 1.f);
    out[i] = acc;
  }
}

// NCHW fp32 -> [N][H][W][C] uint8 : rint(clamp(x, 0, 1) * 255) in fp32 (skimage.img_as_ubyte on float32)
__global__ void to_u8_kernel(const float* __restrict__ in, uint8_t* __restrict__ out, int N, int C, int H, int W) {
  const long long hw = static_cast<long long>(H) * W;
  const long long total = static_cast<long long>(N) * C * hw;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = int(i % C);
    const long long p = (i / C) % hw;
    const long long n = i / (C * hw);
    float v = __ldg(in + (n * C + c) * hw + p);
    v = fminf(fmaxf(v, 0.f), 1.f);
    out[i] = static_cast<uint8_t>(__float2int_rn(__fmul_rn(v, 255.f)));
  }
}

// Y of MATLAB's rgb2ycbcr on a uint8 pixel (utils/util_image.py:129-153): the reference rounds (half to even)
// np.dot([r g b], [65.481 128.553 24.966] / 255) + 16, and numpy's dot is OpenBLAS' dgemv = the fused chain
// fma(b, c2, fma(g, c1, r * c0)) on x86 — about one pixel in 250 000 is an exact tie, so the chain must be the same.
__device__ __forceinline__ int y_of_rgb_u8(int r, int g, int b) {
  const double c0 = 65.481 / 255.0, c1 = 128.553 / 255.0, c2 = 24.966 / 255.0;
  double s = __dmul_rn(double(r), c0);
  s = __fma_rn(double(g), c1, s);
  s = __fma_rn(double(b), c2, s);
  s = __dadd_rn(s, 16.0);
  return int(rint(s));
}

// sum of squared differences over the border-cropped region; acc[0] += SSD (exact, unsigned 64-bit)
__global__ void psnr_ssd_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int H, int W, int C,
                                int border, int ycbcr, unsigned long long* __restrict__ acc) {
  const int h2 = H - 2 * border, w2 = W - 2 * border;
  const long long total = static_cast<long long>(h2) * w2;
  unsigned long long local = 0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int y = int(i / w2) + border, x = int(i % w2) + border;
    const uint8_t* pa = a + (static_cast<long long>(y) * W + x) * C;
    const uint8_t* pb = b + (static_cast<long long>(y) * W + x) * C;
    if (ycbcr) {
      const int d = y_of_rgb_u8(pa[0], pa[1], pa[2]) - y_of_rgb_u8(pb[0], pb[1], pb[2]);
      local += static_cast<unsigned long long>(d * d);
    } else {
      for (int c = 0; c < C; ++c) {
        const int d = int(pa[c]) - int(pb[c]);
        local += static_cast<unsigned long long>(d * d);
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(acc, local);
}

__constant__ double c_ssim_win[121];

// SSIM map of one channel over the "valid" region (H-10) x (W-10) of the border-cropped image; acc[0] += sum (fp64)
__global__ void ssim_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int H, int W, int C, int border,
                            int ycbcr, int chan, double* __restrict__ acc) {
  const int h2 = H - 2 * border, w2 = W - 2 * border;
  const int vh = h2 - 10, vw = w2 - 10;
  const long long total = static_cast<long long>(vh) * vw;
  const double C1 = (0.01 * 255) * (0.01 * 255), C2 = (0.03 * 255) * (0.03 * 255);
  double local = 0.0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int y0 = int(i / vw) + border, x0 = int(i % vw) + border;
    double m1 = 0, m2 = 0, s11 = 0, s22 = 0, s12 = 0;
    for (int dy = 0; dy < 11; ++dy) {
      for (int dx = 0; dx < 11; ++dx) {
        const long long off = (static_cast<long long>(y0 + dy) * W + (x0 + dx)) * C;
        double va, vb;
        if (ycbcr) {
          va = y_of_rgb_u8(a[off], a[off + 1], a[off + 2]);
          vb = y_of_rgb_u8(b[off], b[off + 1], b[off + 2]);
        } else {
          va = a[off + chan], vb = b[off + chan];
        }
        const double wv = c_ssim_win[dy * 11 + dx];
        m1 += wv * va, m2 += wv * vb;
        s11 += wv * va * va, s22 += wv * vb * vb, s12 += wv * va * vb;
      }
    }
    const double m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
    const double v1 = s11 - m11, v2 = s22 - m22, v12 = s12 - m12;
    local += ((2 * m12 + C1) * (2 * v12 + C2)) / ((m11 + m22 + C1) * (v1 + v2 + C2));
  }
  for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(acc, local);
}

int grid_of(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  return int(std::max(1LL, std::min(b, 148LL * 16)));
}
}  // namespace
}  // namespace vk

using namespace vk;
#define VK_ST(s) reinterpret_cast<cudaStream_t>(s)
#define VK_DONE()                                           \
  g_launch_count.fetch_add(1, std::memory_order_relaxed);   \
  return int(cudaGetLastError())

extern "C" int vk_aug8(const float* in, float* out_a, float* out_b, int64_t planes, int32_t h, int32_t w, void* stream) {
  if (!in || !out_a || !out_b || planes <= 0 || h <= 0 || w <= 0) return VK_E_BADARG;
  aug8_kernel<<<grid_of(planes * h * w * 8, 256), 256, 0, VK_ST(stream)>>>(in, out_a, out_b, planes, h, w);
  VK_DONE();
}

extern "C" int vk_aug8_merge(const float* in_a, const float* in_b, float* out, int64_t planes, int32_t h, int32_t w,
                             int32_t clip01, void* stream) {
  if (!in_a || !in_b || !out || planes <= 0 || h <= 0 || w <= 0) return VK_E_BADARG;
  aug8_merge_kernel<<<grid_of(planes * h * w, 256), 256, 0, VK_ST(stream)>>>(in_a, in_b, out, planes, h, w, clip01);
  VK_DONE();
}

extern "C" int vk_to_u8(const float* in, uint8_t* out, int32_t n, int32_t c, int32_t h, int32_t w, void* stream) {
  if (!in || !out || n <= 0 || c <= 0 || h <= 0 || w <= 0) return VK_E_BADARG;
  to_u8_kernel<<<grid_of(static_cast<long long>(n) * c * h * w, 256), 256, 0, VK_ST(stream)>>>(in, out, n, c, h, w);
  VK_DONE();
}

extern "C" int vk_psnr_u8(const uint8_t* a, const uint8_t* b, int32_t h, int32_t w, int32_t c, int32_t border,
                          int32_t ycbcr, uint64_t* ssd_out, void* stream) {
  if (!a || !b || !ssd_out || h <= 0 || w <= 0 || c <= 0 || border < 0) return VK_E_BADARG;
  if (h - 2 * border <= 0 || w - 2 * border <= 0 || (ycbcr && c != 3)) return VK_E_BADARG;
  cudaError_t e = cudaMemsetAsync(ssd_out, 0, sizeof(uint64_t), VK_ST(stream));
  if (e != cudaSuccess) return int(e);
  const long long total = static_cast<long long>(h - 2 * border) * (w - 2 * border);
  psnr_ssd_kernel<<<grid_of(total, 256), 256, 0, VK_ST(stream)>>>(a, b, h, w, c, border, ycbcr,
                                                                 reinterpret_cast<unsigned long long*>(ssd_out));
  VK_DONE();
}

extern "C" int vk_ssim_u8(const uint8_t* a, const uint8_t* b, int32_t h, int32_t w, int32_t c, int32_t border,
                          int32_t ycbcr, const double* window121_host, double* sums_out, void* stream) {
  if (!a || !b || !sums_out || !window121_host || h <= 0 || w <= 0 || c <= 0 || border < 0) return VK_E_BADARG;
  if (h - 2 * border < 11 || w - 2 * border < 11 || (ycbcr && c != 3) || c > 4) return VK_E_BADARG;
  cudaStream_t st = VK_ST(stream);
  cudaError_t e = cudaMemcpyToSymbolAsync(c_ssim_win, window121_host, 121 * sizeof(double), 0, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return int(e);
  const int nch = ycbcr ? 1 : c;
  e = cudaMemsetAsync(sums_out, 0, nch * sizeof(double), st);
  if (e != cudaSuccess) return int(e);
  const long long total = static_cast<long long>(h - 2 * border - 10) * (w - 2 * border - 10);
  for (int ch = 0; ch < nch; ++ch) {
    ssim_kernel<<<grid_of(total, 128), 128, 0, st>>>(a, b, h, w, c, border, ycbcr, ch, sums_out + ch);
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
  }
  return int(cudaGetLastError());
}
