// vk_elementwise.cu — the HBM-bound kernels of the VIRNet hot path (sm_100a):
// layout packing at the NCHW-fp32 boundary, the fused ELBO loss (forward + gradient in
// one pass), the variance-head chain rule, weight packing, gradient-norm + clip + Adam.
// All are pure streaming kernels: coalesced, vectorised where the layout allows, grids
// sized in multiples of the SM count.
#include <algorithm>
#include <cstdio>

#include "../../include/virnet_b200.h"
#include "vk_common.cuh"
#include "vk_host.h"

namespace vk {

constexpr int kSMs = 148;

__device__ __forceinline__ int reflect_idx(int i, int n) {   // F.pad(mode='reflect') on the far side only
  return i < n ? i : 2 * (n - 1) - i;
}
template <typename DT>
__device__ __forceinline__ DT cvt_out(float v);
template <>
__device__ __forceinline__ float cvt_out<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 cvt_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// One NHWC pixel row of `ld` channels, written with 16-byte stores (ld * sizeof(DT) is a multiple of 32);
// f(c) yields channel c as float.
template <typename DT, typename F>
__device__ __forceinline__ void store_row_vec(DT* __restrict__ o, int ld, F f) {
  constexpr int V = 16 / int(sizeof(DT));
  for (int c0 = 0; c0 < ld; c0 += V) {
    uint4 q;
    if constexpr (sizeof(DT) == 4) {
      q = make_uint4(__float_as_uint(f(c0)), __float_as_uint(f(c0 + 1)), __float_as_uint(f(c0 + 2)),
                     __float_as_uint(f(c0 + 3)));
    } else {
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 h2 = __floats2bfloat162_rn(f(c0 + 2 * i), f(c0 + 2 * i + 1));
        w[i] = *reinterpret_cast<uint32_t*>(&h2);
      }
      q = make_uint4(w[0], w[1], w[2], w[3]);
    }
    *reinterpret_cast<uint4*>(o + c0) = q;
  }
}

// ---------------------------------------------------------------------------
// pack_input: NCHW fp32 image (+ conditioning maps) -> NHWC DT, reflect-padded
//   out[n][y][x][c] = c <  C      : img[n][c][ry/sf][rx/sf]
//                     c <  C + E  : f(extra)          (map, or per-sample constant)
//                     else        : 0
// (utils/util_net.py:20-25 pad_input, networks/AttResUNet.py:147-153 cat,
//  networks/VIRNet.py:44 sqrt, :83-95 nearest upsample / repeat)
// ---------------------------------------------------------------------------
template <typename DT>
__global__ void pack_input_kernel(const float* __restrict__ img, int C, int h, int w, int sf,
                                  const float* __restrict__ extra, int E, int extra_is_map, int extra_sqrt,
                                  int eh, int ew, int esf, DT* __restrict__ out, int N, int Hp, int Wp, int ld) {
  const long long total = static_cast<long long>(N) * Hp * Wp;
  const int H = h * sf, W = w * sf;
  for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < total;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = int(p % Wp);
    const int y = int((p / Wp) % Hp);
    const int n = int(p / (static_cast<long long>(Wp) * Hp));
    const int ry = reflect_idx(y, H), rx = reflect_idx(x, W);
    const int iy = ry / sf, ix = rx / sf, ey = ry / esf, ex = rx / esf;
    store_row_vec<DT>(out + p * ld, ld, [&](int c) -> float {
      if (c < C) return __ldg(img + ((static_cast<long long>(n) * C + c) * h + iy) * w + ix);
      const int e = c - C;
      if (e >= E) return 0.f;
      float v = extra_is_map ? __ldg(extra + ((static_cast<long long>(n) * E + e) * eh + ey) * ew + ex)
                             : __ldg(extra + static_cast<long long>(n) * E + e);
      if (extra_sqrt & (1 << e)) v = sqrtf(v);
      return v;
    });
  }
}

// ---------------------------------------------------------------------------
// pack_grad: NCHW fp32 [N][C][h][w] -> NHWC DT [N][Hp][Wp][ld], zero outside (crop backward)
// ---------------------------------------------------------------------------
template <typename DT>
__global__ void pack_grad_kernel(const float* __restrict__ g, int C, int h, int w, DT* __restrict__ out, int N,
                                 int Hp, int Wp, int ld) {
  const long long total = static_cast<long long>(N) * Hp * Wp;
  for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < total;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = int(p % Wp);
    const int y = int((p / Wp) % Hp);
    const int n = int(p / (static_cast<long long>(Wp) * Hp));
    const bool in = (y < h) && (x < w);
    store_row_vec<DT>(out + p * ld, ld, [&](int c) -> float {
      return (in && c < C) ? __ldg(g + ((static_cast<long long>(n) * C + c) * h + y) * w + x) : 0.f;
    });
  }
}

// ---------------------------------------------------------------------------
// sigma_head_bwd: chain rule through  sigma = exp(clamp(l)),  s = sqrt(sigma)
//   g_l = [lo <= l <= hi] * sigma * (g_sigma + g_s / (2 sqrt(sigma)))
// g_s arrives as channel `chan` of the head-conv input gradient (NHWC DT, padded grid);
// the reflect padding folds the mirrored rows/cols back onto their source pixel.
// Output: NHWC DT [N][h][w][ld], channels [0, SC) hold g_l, the rest zero.
// (networks/VIRNet.py:43-44; autograd of torch.exp/clamp/sqrt/F.pad)
// ---------------------------------------------------------------------------
template <typename DT>
__global__ void sigma_head_bwd_kernel(const float* __restrict__ sigma, const float* __restrict__ g_sigma,
                                      const DT* __restrict__ g_in, int ld_in, int chan, int Hp, int Wp,
                                      DT* __restrict__ out, int ld, int N, int SC, int h, int w, float sig_lo,
                                      float sig_hi) {
  const long long total = static_cast<long long>(N) * h * w;
  for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < total;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = int(p % w);
    const int y = int((p / w) % h);
    const int n = int(p / (static_cast<long long>(w) * h));
    // mirrored partners of (y, x) inside the padded grid
    const int y2 = 2 * (h - 1) - y, x2 = 2 * (w - 1) - x;
    const bool my = (y2 >= h) && (y2 < Hp), mx = (x2 >= w) && (x2 < Wp);
    store_row_vec<DT>(out + p * ld, ld, [&](int c) -> float {
      float v = 0.f;
      if (c < SC) {
        const long long si = ((static_cast<long long>(n) * SC + c) * h + y) * w + x;
        const float sg = __ldg(sigma + si);
        float gs = 0.f;
        if (g_in != nullptr) {
          const DT* gi = g_in + static_cast<long long>(n) * Hp * Wp * ld_in + chan + c;
          gs = float(gi[(static_cast<long long>(y) * Wp + x) * ld_in]);
          if (my) gs += float(gi[(static_cast<long long>(y2) * Wp + x) * ld_in]);
          if (mx) gs += float(gi[(static_cast<long long>(y) * Wp + x2) * ld_in]);
          if (my && mx) gs += float(gi[(static_cast<long long>(y2) * Wp + x2) * ld_in]);
        }
        const float gsig = g_sigma != nullptr ? __ldg(g_sigma + si) : 0.f;
        const bool in_range = (sg > sig_lo) && (sg < sig_hi);
        v = in_range ? sg * (gsig + gs * 0.5f * rsqrtf(sg)) : 0.f;
      }
      return v;
    });
  }
}

// ---------------------------------------------------------------------------
// Fused ELBO (loss/ELBO_simple.py:12-53): one pass computes the three means and
// the gradients w.r.t. mu and sigma.  acc[0..2] (double) += lh, kl_gauss, kl_ig sums.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void elbo_denoise_kernel(const float* __restrict__ mu, const float* __restrict__ sigma,
                                    const float* __restrict__ noisy, const float* __restrict__ gt,
                                    const float* __restrict__ beta0, float beta0_scale, int N, int C, int SC,
                                    long long HW, float eps2, float alpha0, float digamma_am1, float gscale,
                                    float* __restrict__ d_mu, float* __restrict__ d_sigma,
                                    double* __restrict__ acc) {
  // one thread per (n, pixel); loops over the C image channels so the 1-channel sigma
  // gradient is produced without atomics
  const long long total = static_cast<long long>(N) * HW;
  const float am1 = alpha0 - 1.f;
  const float inv_m3 = 1.f / (static_cast<float>(N) * C * HW);
  const float inv_ms = 1.f / (static_cast<float>(N) * SC * HW);
  const float half_log2pi = 0.9189385332046727f;
  float s_lh = 0.f, s_kg = 0.f, s_ig = 0.f;
  for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < total;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = p / HW, px = p - n * HW;
    float dsig_acc = 0.f, d_ig = 0.f;
    float inv_beta = 0.f, lbeta = 0.f;
    for (int c = 0; c < C; ++c) {
      const long long i3 = (n * C + c) * HW + px;
      const long long i1 = (n * SC + (SC == 1 ? 0 : c)) * HW + px;
      if (SC != 1 || c == 0) {
        const float sg = __ldg(sigma + i1);
        const float b0 = __ldg(beta0 + i1) * beta0_scale;
        const float beta = sg * alpha0;
        inv_beta = 1.f / beta;
        lbeta = logf(beta);
        s_ig += am1 * (b0 * inv_beta - 1.f) + am1 * (lbeta - logf(b0));
        d_ig = inv_ms * am1 * (inv_beta - b0 * inv_beta * inv_beta) * alpha0;
        dsig_acc = d_ig;
      }
      const float m = __ldg(mu + i3), y = __ldg(noisy + i3), g = __ldg(gt + i3);
      const float e = m - g, r = y - m;
      s_kg += 0.5f * e * e / eps2;
      const float q = r * r + eps2;
      s_lh += 0.5f * (lbeta - digamma_am1 + am1 * inv_beta * q) + half_log2pi;
      if (d_mu != nullptr) d_mu[i3] = gscale * inv_m3 * (e / eps2 - am1 * inv_beta * r);
      const float dl = inv_m3 * 0.5f * (inv_beta - am1 * inv_beta * inv_beta * q) * alpha0;
      if (SC == 1) {
        dsig_acc += dl;
      } else if (d_sigma != nullptr) {
        d_sigma[i1] = gscale * (d_ig + dl);
      }
    }
    if (SC == 1 && d_sigma != nullptr) d_sigma[n * HW + px] = gscale * dsig_acc;
  }
  __shared__ float red[3][32];
  s_lh = warp_sum(s_lh), s_kg = warp_sum(s_kg), s_ig = warp_sum(s_ig);
  const int wid = threadIdx.x >> 5, lid = threadIdx.x & 31;
  if (lid == 0) red[0][wid] = s_lh, red[1][wid] = s_kg, red[2][wid] = s_ig;
  __syncthreads();
  if (wid == 0) {
    const int nw = blockDim.x >> 5;
    float a = lid < nw ? red[0][lid] : 0.f, b = lid < nw ? red[1][lid] : 0.f, c = lid < nw ? red[2][lid] : 0.f;
    a = warp_sum(a), b = warp_sum(b), c = warp_sum(c);
    // one slot per block, no atomics: the sums below are taken in a fixed order, so the loss is bit-reproducible
    if (lid == 0) {
      acc[3 * blockIdx.x + 0] = static_cast<double>(a);
      acc[3 * blockIdx.x + 1] = static_cast<double>(b);
      acc[3 * blockIdx.x + 2] = static_cast<double>(c);
    }
  }
}

// Fixed-order sum of `n` doubles spaced `stride` apart by one 256-thread block: thread t adds elements t, t + 256, ...
// in that order, then a shared-memory tree combines the 256 partial sums.  Every thread returns the total.
__device__ __forceinline__ double block_ordered_sum(const double* __restrict__ v, int n, int stride, double* sm) {
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += v[static_cast<long long>(i) * stride];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (int(threadIdx.x) < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  const double total = sm[0];
  __syncthreads();
  return total;
}

__global__ void __launch_bounds__(256)
elbo_finalize_kernel(const double* __restrict__ acc, int nblocks, double m3, double ms, float* __restrict__ out) {
  __shared__ double sm[256];
  const double lh = block_ordered_sum(acc + 0, nblocks, 3, sm) / m3;
  const double kg = block_ordered_sum(acc + 1, nblocks, 3, sm) / m3;
  const double ig = block_ordered_sum(acc + 2, nblocks, 3, sm) / ms;
  if (threadIdx.x == 0) {
    out[0] = static_cast<float>(lh + kg + ig);
    out[1] = static_cast<float>(lh);
    out[2] = static_cast<float>(kg);
    out[3] = static_cast<float>(ig);
  }
}

// ---------------------------------------------------------------------------
// Weight packing: parameter layout (fp32) -> K-major GEMM operand [taps][rows][ld] (DT)
//   mode 0  conv fprop    : src [Co][Ci][T]      dst[t][co][ci]
//   mode 1  conv dgrad    : src [Co][Ci][T]      dst[T-1-t][ci][co]     (180-degree rotation, Ci/Co swapped)
//   mode 2  convT fprop   : src [Ci][Co][4]      dst[0][t*Co+co][ci]
//   mode 3  convT dgrad   : src [Ci][Co][4]      dst[t][ci][co]         (a 2x2 stride-2 conv)
//   mode 4  s2-conv dgrad : src [Co][Ci][T]      dst[t][ci][co]         (transposed, NOT rotated)
// ---------------------------------------------------------------------------
struct PackDesc {
  const float* src;
  void* dst;
  int dim0, dim1, taps;   // src dims [dim0][dim1][taps]
  int rows, ld;           // dst rows per tap, dst row pitch
  int dst_taps;
  int mode;
  int pad;
};

template <typename DT>
__global__ void pack_weights_kernel(const PackDesc* __restrict__ descs, int round_tf32_flag) {
  const PackDesc d = descs[blockIdx.y];
  // one thread per (row, k) element, looping over the destination taps: the taps of one (Co, Ci) pair are contiguous
  // in the OIHW source (36 B for a 3x3 filter), so every fetched sector is fully used; writes are coalesced along k
  const unsigned rk = unsigned(d.rows) * unsigned(d.ld);
  DT* dst = reinterpret_cast<DT*>(d.dst);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < rk; i += gridDim.x * blockDim.x) {
    const int k = int(i % unsigned(d.ld));
    const int r = int(i / unsigned(d.ld));
    for (int t = 0; t < d.dst_taps; ++t) {
      float v = 0.f;
      int i0 = -1, i1 = -1, it = 0;
      switch (d.mode) {
        case 0: i0 = r, i1 = k, it = t; break;
        case 1: i0 = k, i1 = r, it = d.taps - 1 - t; break;
        case 2: i0 = k, i1 = r % d.dim1, it = r / d.dim1; break;
        case 3: i0 = r, i1 = k, it = t; break;
        default: i0 = k, i1 = r, it = t; break;   // mode 4: transpose Ci/Co without rotation
      }
      if (i0 >= 0 && i0 < d.dim0 && i1 >= 0 && i1 < d.dim1 && it < d.taps)
        v = __ldg(d.src + (static_cast<long long>(i0) * d.dim1 + i1) * d.taps + it);
      if (round_tf32_flag) v = round_tf32(v);
      dst[static_cast<long long>(t) * rk + i] = cvt_out<DT>(v);
    }
  }
}

// ---------------------------------------------------------------------------
// channel_sum: out[c] += sum over pixels of x[p][c]  (ConvTranspose2d bias gradient)
// ---------------------------------------------------------------------------
template <typename DT>
__global__ void channel_sum_kernel(const DT* __restrict__ x, long long npix, int ld, int C, float* __restrict__ out,
                                   float* __restrict__ ws) {
  // block = 256 threads = 8 pixel lanes x 32 channel lanes; blockIdx.y = sample (batched form: per-sample sums)
  const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
  x += static_cast<long long>(blockIdx.y) * npix * ld;
  out += static_cast<long long>(blockIdx.y) * C;
  __shared__ float part[8][33];
  for (int c0 = 0; c0 < C; c0 += 32) {
    const int c = c0 + cl;
    float s = 0.f;
    if (c < C)
      for (long long p = blockIdx.x * 8ll + pl; p < npix; p += gridDim.x * 8ll) s += float(x[p * ld + c]);
    part[pl][cl] = s;
    __syncthreads();
    if (pl == 0 && c < C) {
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) t += part[j][cl];
      if (ws != nullptr) ws[static_cast<long long>(blockIdx.x) * C + c] = t;   // summed in block order by the finalize pass
      else atomicAdd(out + c, t);
    }
    __syncthreads();
  }
}

// 16-byte-vector form for ld % (16 / sizeof(DT)) == 0: a thread keeps the sums of one vector of channels over its
// pixels in registers; lanes are reduced through shared memory, one atomicAdd per (block, channel).
template <typename DT>
__global__ void __launch_bounds__(256)
channel_sum_vec_kernel(const DT* __restrict__ x, long long npix, int ld, int C, float* __restrict__ out,
                       float* __restrict__ ws) {
  constexpr int V = 16 / int(sizeof(DT));
  extern __shared__ float cs_sm[];           // [lanes][ld]
  x += static_cast<long long>(blockIdx.y) * npix * ld;
  out += static_cast<long long>(blockIdx.y) * C;
  const int groups = ld / V, lanes = 256 / groups;
  const int cg = threadIdx.x % groups, pl = threadIdx.x / groups;
  float s[V];
#pragma unroll
  for (int i = 0; i < V; ++i) s[i] = 0.f;
  auto acc = [&](const uint4& r) {
    if constexpr (sizeof(DT) == 2) {
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        s[2 * i] += f.x, s[2 * i + 1] += f.y;
      }
    } else {
      const float* f = reinterpret_cast<const float*>(&r);
#pragma unroll
      for (int i = 0; i < V; ++i) s[i] += f[i];
    }
  };
  if (pl < lanes) {
    // four independent 16-byte loads in flight per thread (narrow tensors are latency-bound otherwise)
    const long long step = static_cast<long long>(gridDim.x) * lanes;
    long long p = blockIdx.x * static_cast<long long>(lanes) + pl;
    for (; p + 3 * step < npix; p += 4 * step) {
      const uint4 r0 = *reinterpret_cast<const uint4*>(x + p * ld + cg * V);
      const uint4 r1 = *reinterpret_cast<const uint4*>(x + (p + step) * ld + cg * V);
      const uint4 r2 = *reinterpret_cast<const uint4*>(x + (p + 2 * step) * ld + cg * V);
      const uint4 r3 = *reinterpret_cast<const uint4*>(x + (p + 3 * step) * ld + cg * V);
      acc(r0), acc(r1), acc(r2), acc(r3);
    }
    for (; p < npix; p += step) acc(*reinterpret_cast<const uint4*>(x + p * ld + cg * V));
  }
  if (pl < lanes) {
#pragma unroll
    for (int i = 0; i < V; ++i) cs_sm[pl * ld + cg * V + i] = s[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float t = 0.f;
    for (int l = 0; l < lanes; ++l) t += cs_sm[l * ld + c];
    if (ws != nullptr) ws[static_cast<long long>(blockIdx.x) * C + c] = t;
    else atomicAdd(out + c, t);
  }
}

// out[c] += sum over the nblocks per-block partial sums, in block order (no atomics: bit-reproducible)
// block = 32 channels x 8 lanes; lane l adds the partials of blocks l, l + 8, ... in that order, then the 8 lane sums are
// added in lane order: a fixed summation tree
__global__ void __launch_bounds__(256)
channel_sum_finalize_kernel(const float* __restrict__ ws, int nblocks, int C, float* __restrict__ out) {
  const int cl = threadIdx.x & 31, l = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  __shared__ float part[8][33];
  float t = 0.f;
  if (c < C)
    for (int b = l; b < nblocks; b += 8) t += ws[static_cast<long long>(b) * C + c];
  part[l][cl] = t;
  __syncthreads();
  if (l == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += part[j][cl];
    out[c] += s;
  }
}

// ---------------------------------------------------------------------------
// Gradient norm -> clip -> Adam over flat fp32 buffers, grouped by sub-network
// (train_denoising_syn.py:182-184: clip_grad_norm_ per sub-net, optim.Adam.step)
// ---------------------------------------------------------------------------
struct AdamGroup {
  long long begin, end;   // element range in the flat buffers
  float max_norm;
  int pad;
};

__global__ void grad_sqnorm_kernel(const float* __restrict__ g, const AdamGroup* __restrict__ groups, int ngroups,
                                   float gscale, double* __restrict__ sq) {
  const int gi = blockIdx.y;
  const AdamGroup gr = groups[gi];
  float s = 0.f;
  for (long long i = gr.begin + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < gr.end;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = g[i] * gscale;
    s += v * v;
  }
  __shared__ float red[32];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float a = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    a = warp_sum(a);
    if (threadIdx.x == 0) sq[static_cast<long long>(gi) * gridDim.x + blockIdx.x] = static_cast<double>(a);   // per-block slot
  }
}

__global__ void adam_clip_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, const AdamGroup* __restrict__ groups, int ngroups,
                                 const double* __restrict__ sq, float gscale, float lr, float beta1, float beta2,
                                 float eps, float bc1, float bc2_sqrt, float* __restrict__ norms_out,
                                 const float* __restrict__ hyper) {
  if (hyper != nullptr) lr = hyper[0], bc1 = hyper[1], bc2_sqrt = hyper[2];     // CUDA-graph replay: per-step values
  const int gi = blockIdx.y;
  const AdamGroup gr = groups[gi];
  // squared norm of the group = the per-block partials of grad_sqnorm_kernel (same grid) summed in a fixed order by
  // every block: no atomics, so the clip coefficient and with it the whole update are bit-reproducible
  __shared__ double sq_sm[256];
  const float total = static_cast<float>(sqrt(block_ordered_sum(sq + static_cast<long long>(gi) * gridDim.x, int(gridDim.x), 1, sq_sm)));
  const float coef = fminf(gr.max_norm / (total + 1e-6f), 1.f) * gscale;   // clip_grad_norm_ semantics
  if (norms_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) norms_out[gi] = total;
  const float step = lr / bc1;
  auto upd = [&](float gv, float& pv, float& mv, float& vv_) {
    const float gg = gv * coef;
    const float mm = beta1 * mv + (1.f - beta1) * gg;
    const float vv = beta2 * vv_ + (1.f - beta2) * gg * gg;
    mv = mm, vv_ = vv;
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pv -= step * mm / denom;
  };
  if (((gr.begin | gr.end) & 3) == 0) {
    // every tensor of the flat buffers is padded to 4 elements: 16-byte vectors (7 streams, 28 B per parameter)
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* p4 = reinterpret_cast<float4*>(p);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (long long i = (gr.begin >> 2) + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < (gr.end >> 2);
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      const float4 gv = g4[i];
      float4 pv = p4[i], mv = m4[i], vv = v4[i];
      upd(gv.x, pv.x, mv.x, vv.x), upd(gv.y, pv.y, mv.y, vv.y), upd(gv.z, pv.z, mv.z, vv.z), upd(gv.w, pv.w, mv.w, vv.w);
      m4[i] = mv, v4[i] = vv, p4[i] = pv;
    }
    return;
  }
  for (long long i = gr.begin + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < gr.end;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float pv = p[i], mv = m[i], vv = v[i];
    upd(g[i], pv, mv, vv);
    m[i] = mv, v[i] = vv, p[i] = pv;
  }
}

static inline int grid_for(long long total, int threads, int max_waves = 16) {
  long long b = (total + threads - 1) / threads;
  const long long cap = static_cast<long long>(kSMs) * max_waves;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return int(b);
}

}  // namespace vk

using namespace vk;
#define VK_ST(s) reinterpret_cast<cudaStream_t>(s)
#define VK_LAUNCHED()                                        \
  g_launch_count.fetch_add(1, std::memory_order_relaxed);    \
  return int(cudaGetLastError())

extern "C" int vk_pack_input(int32_t dtype, const float* img, int32_t n, int32_t c, int32_t h, int32_t w, int32_t sf,
                             const float* extra, int32_t e, int32_t extra_is_map, int32_t extra_sqrt_mask,
                             int32_t eh, int32_t ew, int32_t esf, void* out, int32_t hp, int32_t wp, int32_t ld,
                             void* stream) {
  if (img == nullptr || out == nullptr || n <= 0 || c <= 0 || c + e > ld || sf <= 0) return VK_E_BADARG;
  if (e > 0 && extra == nullptr) return VK_E_BADARG;
  if (hp < h * sf || wp < w * sf || hp > 2 * h * sf - 1 || wp > 2 * w * sf - 1) return VK_E_BADARG;
  if (esf <= 0) esf = 1;
  const long long total = static_cast<long long>(n) * hp * wp;
  const int grid = grid_for(total, 256);
  if (dtype == VK_BF16)
    pack_input_kernel<__nv_bfloat16><<<grid, 256, 0, VK_ST(stream)>>>(
        img, c, h, w, sf, extra, e, extra_is_map, extra_sqrt_mask, eh, ew, esf,
        reinterpret_cast<__nv_bfloat16*>(out), n, hp, wp, ld);
  else if (dtype == VK_TF32)
    pack_input_kernel<float><<<grid, 256, 0, VK_ST(stream)>>>(img, c, h, w, sf, extra, e, extra_is_map,
                                                             extra_sqrt_mask, eh, ew, esf,
                                                             reinterpret_cast<float*>(out), n, hp, wp, ld);
  else
    return VK_E_BADARG;
  VK_LAUNCHED();
}

extern "C" int vk_pack_grad(int32_t dtype, const float* g, int32_t n, int32_t c, int32_t h, int32_t w, void* out,
                            int32_t hp, int32_t wp, int32_t ld, void* stream) {
  if (g == nullptr || out == nullptr || n <= 0 || c <= 0 || c > ld || hp < h || wp < w) return VK_E_BADARG;
  const long long total = static_cast<long long>(n) * hp * wp;
  const int grid = grid_for(total, 256);
  if (dtype == VK_BF16)
    pack_grad_kernel<__nv_bfloat16><<<grid, 256, 0, VK_ST(stream)>>>(g, c, h, w, reinterpret_cast<__nv_bfloat16*>(out),
                                                                    n, hp, wp, ld);
  else if (dtype == VK_TF32)
    pack_grad_kernel<float><<<grid, 256, 0, VK_ST(stream)>>>(g, c, h, w, reinterpret_cast<float*>(out), n, hp, wp, ld);
  else
    return VK_E_BADARG;
  VK_LAUNCHED();
}

extern "C" int vk_sigma_head_bwd(int32_t dtype, const float* sigma, const float* g_sigma, const void* g_in,
                                 int32_t ld_in, int32_t chan, int32_t hp, int32_t wp, void* out, int32_t ld,
                                 int32_t n, int32_t sc, int32_t h, int32_t w, float log_lo, float log_hi,
                                 void* stream) {
  if (sigma == nullptr || out == nullptr || n <= 0 || sc <= 0 || sc > ld) return VK_E_BADARG;
  if (g_in != nullptr && (chan < 0 || chan + sc > ld_in || hp < h || wp < w)) return VK_E_BADARG;
  const long long total = static_cast<long long>(n) * h * w;
  const int grid = grid_for(total, 256);
  const float lo = expf(log_lo), hi = expf(log_hi);
  if (dtype == VK_BF16)
    sigma_head_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, VK_ST(stream)>>>(
        sigma, g_sigma, reinterpret_cast<const __nv_bfloat16*>(g_in), ld_in, chan, hp, wp,
        reinterpret_cast<__nv_bfloat16*>(out), ld, n, sc, h, w, lo, hi);
  else if (dtype == VK_TF32)
    sigma_head_bwd_kernel<float><<<grid, 256, 0, VK_ST(stream)>>>(sigma, g_sigma,
                                                                 reinterpret_cast<const float*>(g_in), ld_in, chan,
                                                                 hp, wp, reinterpret_cast<float*>(out), ld, n, sc,
                                                                 h, w, lo, hi);
  else
    return VK_E_BADARG;
  VK_LAUNCHED();
}

extern "C" int vk_elbo_denoise(const float* mu, const float* sigma, const float* noisy, const float* gt,
                               const float* beta0, float beta0_scale, int32_t n, int32_t c, int32_t sc, int32_t h,
                               int32_t w, float eps2, float alpha0, float digamma_alpha0_m1, float grad_scale, float* d_mu,
                               float* d_sigma, double* acc_ws, int32_t acc_ws_doubles, float* out4, void* stream) {
  if (!mu || !sigma || !noisy || !gt || !beta0 || !acc_ws || !out4) return VK_E_BADARG;
  if (n <= 0 || c <= 0 || (sc != 1 && sc != c)) return VK_E_BADARG;
  cudaStream_t st = VK_ST(stream);
  const long long hw = static_cast<long long>(h) * w;
  const int grid = grid_for(static_cast<long long>(n) * hw, 256, 4);
  static_assert(kSMs * 4 <= VK_REDUCE_MAX_BLOCKS, "scratch sizing");
  if (acc_ws_doubles < 3 * grid) return VK_E_BADARG;
  elbo_denoise_kernel<<<grid, 256, 0, st>>>(mu, sigma, noisy, gt, beta0, beta0_scale, n, c, sc, hw, eps2, alpha0,
                                            digamma_alpha0_m1, grad_scale, d_mu, d_sigma, acc_ws);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  elbo_finalize_kernel<<<1, 256, 0, st>>>(acc_ws, grid, double(n) * c * hw, double(n) * sc * hw, out4);
  VK_LAUNCHED();
}

extern "C" int vk_pack_weights(int32_t dtype, const void* descs_dev, int32_t ndesc, int64_t max_elems,
                               int32_t round_tf32_flag, void* stream) {
  if (descs_dev == nullptr || ndesc <= 0) return VK_E_BADARG;
  dim3 grid(grid_for((max_elems + 8) / 9, 256, 2), ndesc);     // ~one thread per (row, k) of the largest 3x3 layer
  if (dtype == VK_BF16)
    pack_weights_kernel<__nv_bfloat16><<<grid, 256, 0, VK_ST(stream)>>>(
        reinterpret_cast<const PackDesc*>(descs_dev), 0);
  else if (dtype == VK_TF32)
    pack_weights_kernel<float><<<grid, 256, 0, VK_ST(stream)>>>(reinterpret_cast<const PackDesc*>(descs_dev),
                                                               round_tf32_flag);
  else
    return VK_E_BADARG;
  VK_LAUNCHED();
}

extern "C" int vk_channel_sum(int32_t dtype, const void* x, int64_t npix, int32_t ld, int32_t c, float* out,
                              float* ws, int64_t ws_floats, void* stream) {
  if (x == nullptr || out == nullptr || npix <= 0 || c <= 0 || c > ld) return VK_E_BADARG;
  if (dtype != VK_BF16 && dtype != VK_TF32) return VK_E_BADARG;
  if (ws != nullptr && ws_floats < static_cast<int64_t>(VK_REDUCE_MAX_BLOCKS) * c) return VK_E_BADARG;
  auto finalize = [&](int nblocks) -> int {
    if (ws == nullptr) return int(cudaGetLastError());
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    channel_sum_finalize_kernel<<<(c + 31) / 32, 256, 0, VK_ST(stream)>>>(ws, nblocks, c, out);
    return int(cudaGetLastError());
  };
  {
    const int vec = dtype == VK_BF16 ? 8 : 4;
    if ((dtype == VK_BF16 || dtype == VK_TF32) && ld % vec == 0 && ld / vec <= 256 && npix >= 4096) {
      const int lanes = 256 / (ld / vec);
      const size_t smem = size_t(lanes) * ld * sizeof(float);
      if (smem <= 48 * 1024) {
        const int vgrid = int(std::min<long long>((npix + lanes - 1) / lanes, kSMs * 4));
        if (dtype == VK_BF16)
          channel_sum_vec_kernel<__nv_bfloat16><<<vgrid, 256, smem, VK_ST(stream)>>>(
              reinterpret_cast<const __nv_bfloat16*>(x), npix, ld, c, out, ws);
        else
          channel_sum_vec_kernel<float><<<vgrid, 256, smem, VK_ST(stream)>>>(reinterpret_cast<const float*>(x), npix, ld, c,
                                                                            out, ws);
        g_launch_count.fetch_add(1, std::memory_order_relaxed);
        return finalize(vgrid);
      }
    }
  }
  const int grid = int(std::min<long long>((npix + 7) / 8, kSMs * 4));
  if (dtype == VK_BF16)
    channel_sum_kernel<__nv_bfloat16><<<grid, 256, 0, VK_ST(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                                      npix, ld, c, out, ws);
  else
    channel_sum_kernel<float><<<grid, 256, 0, VK_ST(stream)>>>(reinterpret_cast<const float*>(x), npix, ld, c, out, ws);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return finalize(grid);
}

extern "C" int vk_channel_sum_batched(int32_t dtype, const void* x, int32_t n, int64_t npix, int32_t ld, int32_t c,
                                      float* out, void* stream) {
  if (x == nullptr || out == nullptr || n <= 0 || npix <= 0 || c <= 0 || c > ld) return VK_E_BADARG;
  const dim3 grid(unsigned(std::min<long long>((npix + 7) / 8, std::max(1, kSMs * 4 / n))), unsigned(n));
  if (dtype == VK_BF16)
    channel_sum_kernel<__nv_bfloat16><<<grid, 256, 0, VK_ST(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                                      npix, ld, c, out, nullptr);
  else if (dtype == VK_TF32)
    channel_sum_kernel<float><<<grid, 256, 0, VK_ST(stream)>>>(reinterpret_cast<const float*>(x), npix, ld, c, out, nullptr);
  else
    return VK_E_BADARG;
  VK_LAUNCHED();
}

extern "C" int vk_adam_clip_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                                 const void* groups_dev, int32_t ngroups, int64_t max_group_elems, double* sq_ws,
                                 int32_t sq_ws_doubles, float grad_scale, float lr, float beta1, float beta2, float eps,
                                 int32_t step, float* norms_out, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !groups_dev || !sq_ws || ngroups <= 0 || step <= 0)
    return VK_E_BADARG;
  cudaStream_t st = VK_ST(stream);
  const AdamGroup* groups = reinterpret_cast<const AdamGroup*>(groups_dev);
  dim3 grid(grid_for(max_group_elems, 256, 4), ngroups);
  if (static_cast<long long>(sq_ws_doubles) < static_cast<long long>(grid.x) * ngroups) return VK_E_BADARG;
  grad_sqnorm_kernel<<<grid, 256, 0, st>>>(grads, groups, ngroups, grad_scale, sq_ws);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  // bias corrections in double from the fp32 betas — the values a caller of vk_adam_clip_step_dev computes on the host
  // (virnet_b200/trainer.py), so that the eager and the graph-replayed step are bit-identical
  const float bc1 = float(1.0 - pow(double(beta1), double(step)));
  const float bc2_sqrt = float(sqrt(1.0 - pow(double(beta2), double(step))));
  adam_clip_kernel<<<grid, 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, groups, ngroups, sq_ws, grad_scale, lr,
                                         beta1, beta2, eps, bc1, bc2_sqrt, norms_out, nullptr);
  VK_LAUNCHED();
}

extern "C" int vk_adam_clip_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                                     const void* groups_dev, int32_t ngroups, int64_t max_group_elems, double* sq_ws,
                                     int32_t sq_ws_doubles, float grad_scale, float beta1, float beta2, float eps,
                                     const float* hyper_dev, float* norms_out, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !groups_dev || !sq_ws || !hyper_dev || ngroups <= 0) return VK_E_BADARG;
  cudaStream_t st = VK_ST(stream);
  const AdamGroup* groups = reinterpret_cast<const AdamGroup*>(groups_dev);
  dim3 grid(grid_for(max_group_elems, 256, 4), ngroups);
  if (static_cast<long long>(sq_ws_doubles) < static_cast<long long>(grid.x) * ngroups) return VK_E_BADARG;
  grad_sqnorm_kernel<<<grid, 256, 0, st>>>(grads, groups, ngroups, grad_scale, sq_ws);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  adam_clip_kernel<<<grid, 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, groups, ngroups, sq_ws, grad_scale, 0.f, beta1,
                                         beta2, eps, 1.f, 1.f, norms_out, hyper_dev);
  VK_LAUNCHED();
}
