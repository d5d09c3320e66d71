// vk_synth.cu — device-side synthesis of denoising training batches (sm_100a): the arithmetic of
// datasets/DenoisingDatasets.py:180-253 (SimulateTrain.__getitem__) after the patch has been cropped:
//   uint8 HWC patch -> float32 clean image (skimage.img_as_float32: x * (1/255) in fp32)
//   sigma map: Gaussian bump exp(-((i-ch)^2 + (j-cw)^2) / (2 s^2)) min-max normalised onto [down, up] ('niid',
//              :189-201, util_denoising.py:12-22; evaluated in fp64 like numpy) or a constant ('iid', :203-209)
//   noisy = clean + randn * sigma (optionally clipped to [0, 1]), the 8-way flip / rotate augmentation
//   (utils/util_image.py:391-436) applied to all three, sigma_gt = max(sigma^2, 1e-10); NCHW fp32 out.
// The host draws the per-sample scalars (Python `random`, the reference's order) and the normal noise; one thread
// produces one output pixel, so writes are coalesced and every input byte is read once.
#include <algorithm>

#include "../../include/virnet_b200.h"
#include "vk_common.cuh"
#include "vk_host.h"

namespace vk {

// source coordinates (si, sj) of output pixel (i, j) under numpy's flipud / rot90 on a P x P image
__device__ __forceinline__ void aug_source(int mode, int i, int j, int P, int& si, int& sj) {
  switch (mode) {
    case 0: si = i, sj = j; break;
    case 1: si = P - 1 - i, sj = j; break;               // flipud
    case 2: si = j, sj = P - 1 - i; break;               // rot90 (counter-clockwise)
    case 3: si = j, sj = i; break;                       // rot90 then flipud
    case 4: si = P - 1 - i, sj = P - 1 - j; break;       // rot180
    case 5: si = i, sj = P - 1 - j; break;               // rot180 then flipud
    case 6: si = P - 1 - j, sj = i; break;               // rot270
    default: si = P - 1 - j, sj = P - 1 - i; break;      // rot270 then flipud
  }
}

// params[n] = {center_h, center_w, scale, up, down, iid_level}; scale <= 0 selects the constant ('iid') map
__global__ void synth_denoise_kernel(const uint8_t* __restrict__ patches, const double* __restrict__ params,
                                     const int* __restrict__ aug, const float* __restrict__ noise, int P, int C, int clip,
                                     float* __restrict__ im_noisy, float* __restrict__ im_gt,
                                     float* __restrict__ sigma_gt) {
  const int n = blockIdx.y;
  const double* pr = params + n * 6;
  const double ch = pr[0], cw = pr[1], sc = pr[2], up = pr[3], down = pr[4];
  const bool niid = sc > 0.0;
  double kmin = 0.0, kinv = 0.0;
  const double den = 2.0 * (sc * sc);
  if (niid) {
    // extrema of the bump over the integer grid: nearest / farthest pixel to the centre along each axis
    auto near_far = [&](double c, double& dn, double& df) {
      const double lo = fmin(fmax(floor(c), 0.0), double(P - 1)), hi = fmin(fmax(ceil(c), 0.0), double(P - 1));
      const double a = (lo - c) * (lo - c), b = (hi - c) * (hi - c);
      dn = fmin(a, b);
      const double e0 = (0.0 - c) * (0.0 - c), e1 = (double(P - 1) - c) * (double(P - 1) - c);
      df = fmax(e0, e1);
    };
    double nh, fh, nw, fw;
    near_far(ch, nh, fh);
    near_far(cw, nw, fw);
    const double kmax = exp((-nh - nw) / den);
    kmin = exp((-fh - fw) / den);
    kinv = 1.0 / (kmax - kmin);
  }
  const int mode = aug[n];
  const long long plane = static_cast<long long>(P) * P;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < P * P; idx += gridDim.x * blockDim.x) {
    const int i = idx / P, j = idx - i * P;
    int si, sj;
    aug_source(mode, i, j, P, si, sj);
    float sigma;
    if (niid) {
      const double di = double(si) - ch, dj = double(sj) - cw;
      const double kk = exp((-(di * di) - (dj * dj)) / den);
      sigma = float(down + (kk - kmin) * kinv * (up - down));
    } else {
      sigma = float(pr[5]);
    }
    const long long src = (static_cast<long long>(n) * plane + static_cast<long long>(si) * P + sj) * C;
    for (int c = 0; c < C; ++c) {
      const float g = __fmul_rn(float(patches[src + c]), 1.0f / 255.0f);
      float y = __fadd_rn(g, __fmul_rn(noise[src + c], sigma));
      if (clip) y = fminf(fmaxf(y, 0.f), 1.f);
      const long long dst = (static_cast<long long>(n) * C + c) * plane + idx;
      im_gt[dst] = g;
      im_noisy[dst] = y;
    }
    sigma_gt[static_cast<long long>(n) * plane + idx] = fmaxf(__fmul_rn(sigma, sigma), 1e-10f);
  }
}

}  // namespace vk

extern "C" int vk_synth_denoise(const uint8_t* patches, const double* params, const int32_t* aug, const float* noise,
                                int32_t n, int32_t p, int32_t c, int32_t clip, float* im_noisy, float* im_gt,
                                float* sigma_gt, void* stream) {
  if (!patches || !params || !aug || !noise || !im_noisy || !im_gt || !sigma_gt || n <= 0 || p <= 0 || c <= 0)
    return VK_E_BADARG;
  const dim3 grid(unsigned(std::min((p * p + 255) / 256, 64)), unsigned(n));
  vk::synth_denoise_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(patches, params, aug, noise, p, c, clip,
                                                                                   im_noisy, im_gt, sigma_gt);
  vk::g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return int(cudaGetLastError());
}
