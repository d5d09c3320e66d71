// vk_sisr.cu — the small, latency-bound kernels of the super-resolution forward path (sm_100a):
// KernelNet's 9x9 stride-4 head and channel-attention layers (networks/KNet.py:12-59), the global-average
// heads of SNet / KNet (networks/DnCNN.py:30-33,42, networks/KNet.py:55-58, networks/VIRNet.py:81) and
// the SFT modulation MLPs of AttLayer (networks/AttResUNet.py:11-32).  The 3x3 convolutions between
// them run on the tcgen05 kernels of vk_conv_*.  All of these touch a few hundred KB: one CTA per
// sample, everything staged through shared memory / registers, no atomics.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "../../include/virnet_b200.h"
#include "vk_common.cuh"
#include "vk_host.h"

namespace vk {

template <typename DT>
__device__ __forceinline__ float ld_f(const DT* p);
template <>
__device__ __forceinline__ float ld_f<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ld_f<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename DT>
__device__ __forceinline__ void st_f(DT* p, float v);
template <>
__device__ __forceinline__ void st_f<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void st_f<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// ---------------------------------------------------------------------------
// KNet head: Conv2d(c -> cout, k = 9, stride 4, pad 4, no bias), NCHW fp32 in, NHWC DT out.
// One thread per (output pixel, output channel), a block = 256 / ld pixels x ld channels.  Per input channel the
// 81 x cout filter slice is staged TRANSPOSED in shared memory ([tap][co]: the lanes of a warp read consecutive words;
// straight from the OIHW tensor they would touch 32 cache lines per load), the input pixel is a warp-wide broadcast
// through L1 (every input pixel is shared by ~5 windows and all `cout` threads of a pixel).
// ---------------------------------------------------------------------------
template <typename DT>
__global__ void __launch_bounds__(256)
knet_head_kernel(const float* __restrict__ x, const float* __restrict__ w, DT* __restrict__ out, int N, int C, int H,
                 int W, int OH, int OW, int cout, int ld) {
  constexpr int G = 4;                   // pixel groups per staged filter slice
  extern __shared__ float ws[];          // [81][ld]
  const int ppb = 256 / ld;              // pixels per group (host guarantees ld <= 256)
  const int co = threadIdx.x % ld, pl = threadIdx.x / ld;
  const long long npix = static_cast<long long>(N) * OH * OW;
  for (long long p0 = static_cast<long long>(blockIdx.x) * ppb * G; p0 < npix;
       p0 += static_cast<long long>(gridDim.x) * ppb * G) {
    float acc[G];
#pragma unroll
    for (int g = 0; g < G; ++g) acc[g] = 0.f;
    for (int c = 0; c < C; ++c) {
      __syncthreads();                   // the previous slice has been consumed
#pragma unroll 5
      for (int i = threadIdx.x; i < 81 * ld; i += 256) {
        const int t = i % 81, k = i / 81;            // consecutive threads read consecutive taps of one filter
        ws[t * ld + k] = k < cout ? __ldg(w + (static_cast<long long>(k) * C + c) * 81 + t) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const long long pix = p0 + g * ppb + pl;
        if (pl >= ppb || pix >= npix) continue;
        const int ox = int(pix % OW), oy = int((pix / OW) % OH), n = int(pix / (static_cast<long long>(OW) * OH));
        const float* xp = x + (static_cast<long long>(n) * C + c) * H * W;
        float a = acc[g];
        for (int r = 0; r < 9; ++r) {
          const int iy = oy * 4 - 4 + r;
          if (iy < 0 || iy >= H) continue;
          for (int s = 0; s < 9; ++s) {
            const int ix = ox * 4 - 4 + s;
            if (ix < 0 || ix >= W) continue;
            a = fmaf(__ldg(xp + iy * W + ix), ws[(r * 9 + s) * ld + co], a);
          }
        }
        acc[g] = a;
      }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const long long pix = p0 + g * ppb + pl;
      if (pl < ppb && pix < npix) st_f<DT>(out + pix * ld + co, acc[g]);
    }
  }
}

// ---------------------------------------------------------------------------
// CALayer + residual of RB_Layer (networks/KNet.py:23-26, 37-39), one CTA per sample:
//   y = mean_pix f;  z = LReLU(W1 y + b1);  s = sigmoid(W2 z + b2);  out = f * s + skip
// ---------------------------------------------------------------------------
template <typename DT>
__global__ void __launch_bounds__(1024) ca_layer_kernel(const DT* __restrict__ f, const DT* __restrict__ skip, const float* __restrict__ w1,
                                const float* __restrict__ b1, const float* __restrict__ w2,
                                const float* __restrict__ b2, DT* __restrict__ out, int npix, int C, int R, int ld,
                                float alpha) {
  extern __shared__ float sm[];          // [blockDim.x / C][C] partial sums, then y[C], z[R], s[C]
  const int n = blockIdx.x;
  const DT* fp = f + static_cast<long long>(n) * npix * ld;
  const DT* sp = skip + static_cast<long long>(n) * npix * ld;
  DT* op = out + static_cast<long long>(n) * npix * ld;
  const int lanes = blockDim.x / C;      // pixel lanes
  const int c = threadIdx.x % C, pl = threadIdx.x / C;
  float part = 0.f;
  if (pl < lanes)
    for (int p = pl; p < npix; p += lanes) part += ld_f<DT>(fp + static_cast<long long>(p) * ld + c);
  float* partial = sm;
  float* y = sm + lanes * C;
  float* z = y + C;
  float* s = z + R;
  if (pl < lanes) partial[pl * C + c] = part;
  __syncthreads();
  if (threadIdx.x < C) {
    float t = 0.f;
    for (int l = 0; l < lanes; ++l) t += partial[l * C + threadIdx.x];
    y[threadIdx.x] = t / float(npix);
  }
  __syncthreads();
  if (threadIdx.x < R) {
    float t = b1[threadIdx.x];
    for (int k = 0; k < C; ++k) t = fmaf(w1[threadIdx.x * C + k], y[k], t);
    z[threadIdx.x] = t > 0.f ? t : t * alpha;
  }
  __syncthreads();
  if (threadIdx.x < C) {
    float t = b2[threadIdx.x];
    for (int k = 0; k < R; ++k) t = fmaf(w2[threadIdx.x * R + k], z[k], t);
    s[threadIdx.x] = 1.f / (1.f + expf(-t));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < npix * ld; i += blockDim.x) {
    const int cc = i % ld;
    float v = 0.f;
    if (cc < C) v = ld_f<DT>(fp + i) * s[cc] + ld_f<DT>(sp + i);
    st_f<DT>(op + i, v);
  }
}

// ---------------------------------------------------------------------------
// Global average over NCHW fp32 planes followed by a per-channel head:
//   exp(clamp(mean, lo, hi)) for channels in exp_mask, tanh(mean) for channels in tanh_mask.
// SNet with noise_avg (DnCNN.py:30-33 + VIRNet.py:81) and KNet's tail (KNet.py:55-58).
// ---------------------------------------------------------------------------
__global__ void gap_head_kernel(const float* __restrict__ x, int hw, int C, unsigned exp_mask, unsigned tanh_mask,
                                float lo, float hi, float* __restrict__ out) {
  __shared__ float red[32];
  const int plane = blockIdx.x;          // n * C + c
  const int c = plane % C;
  const float* p = x + static_cast<long long>(plane) * hw;
  float t = 0.f;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) t += p[i];
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < int(blockDim.x >> 5); ++i) tot += red[i];
    float m = tot / float(hw);
    if (exp_mask & (1u << c)) m = expf(fminf(fmaxf(m, lo), hi));
    if (tanh_mask & (1u << c)) m = tanhf(m);
    out[plane] = m;
  }
}

// ---------------------------------------------------------------------------
// AttLayer (networks/AttResUNet.py:27-32) on per-sample CONSTANT conditioning maps: the 1x1-conv MLP
// commutes with the spatial repeat, so (mul, add) are per-(sample, channel) scalars:
//   f1 = LReLU(W1 e + b1); f2 = LReLU(W2 f1 + b2); mul = sigmoid(Wm f2 + bm); add = Wa f2 + ba
// e[n][k] = sqrt(extra[n][k]) for k in sqrt_mask (the noise variance), extra[n][k] otherwise.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void sft_mlp_body(int n, const float* __restrict__ extra, int E, unsigned sqrt_mask, const float* __restrict__ w1,
                               const float* __restrict__ b1, int C1, const float* __restrict__ w2,
                               const float* __restrict__ b2, int C2, const float* __restrict__ wm,
                               const float* __restrict__ bm, const float* __restrict__ wa,
                               const float* __restrict__ ba, int C, float alpha, float* __restrict__ mul,
                               float* __restrict__ add) {
  extern __shared__ float sm[];          // e[E], f1[C1], f2[C2]
  float* e = sm;
  float* f1 = e + E;
  float* f2 = f1 + C1;
  if (threadIdx.x < E) {
    float v = extra[n * E + threadIdx.x];
    e[threadIdx.x] = (sqrt_mask & (1u << threadIdx.x)) ? sqrtf(v) : v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C1; i += blockDim.x) {
    float t = b1[i];
    for (int k = 0; k < E; ++k) t = fmaf(w1[i * E + k], e[k], t);
    f1[i] = t > 0.f ? t : t * alpha;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C2; i += blockDim.x) {
    float t = b2[i];
    for (int k = 0; k < C1; ++k) t = fmaf(w2[i * C1 + k], f1[k], t);
    f2[i] = t > 0.f ? t : t * alpha;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    float tm = bm[i], ta = ba[i];
    for (int k = 0; k < C2; ++k) {
      tm = fmaf(wm[i * C2 + k], f2[k], tm);
      ta = fmaf(wa[i * C2 + k], f2[k], ta);
    }
    mul[n * C + i] = 1.f / (1.f + expf(-tm));
    add[n * C + i] = ta;
  }
}

__global__ void sft_mlp_kernel(const float* __restrict__ extra, int E, unsigned sqrt_mask, const float* __restrict__ w1,
                               const float* __restrict__ b1, int C1, const float* __restrict__ w2,
                               const float* __restrict__ b2, int C2, const float* __restrict__ wm,
                               const float* __restrict__ bm, const float* __restrict__ wa,
                               const float* __restrict__ ba, int C, float alpha, float* __restrict__ mul,
                               float* __restrict__ add) {
  sft_mlp_body(blockIdx.x, extra, E, sqrt_mask, w1, b1, C1, w2, b2, C2, wm, bm, wa, ba, C, alpha, mul, add);
}

// every AttLayer of the network in one launch: blockIdx.y selects the layer's descriptor (mirror of vk_sft_desc)
struct SftDesc {
  const float *w1, *b1, *w2, *b2, *wm, *bm, *wa, *ba;
  float *gw1, *gb1, *gw2, *gb2, *gwm, *gbm, *gwa, *gba;
  float *mul, *add, *dmul, *dadd;
  int c1, c2, c, pad;
};

__global__ void sft_mlp_batched_kernel(const SftDesc* __restrict__ descs, const float* __restrict__ extra, int E,
                                       unsigned sqrt_mask, float alpha) {
  const SftDesc d = descs[blockIdx.y];
  sft_mlp_body(blockIdx.x, extra, E, sqrt_mask, d.w1, d.b1, d.c1, d.w2, d.b2, d.c2, d.wm, d.bm, d.wa, d.ba, d.c, alpha,
               d.mul, d.add);
}

// F.interpolate(x, scale_factor=sf, mode='nearest') on NCHW fp32 (networks/VIRNet.py:83): the global residual of RNet
__global__ void upsample_nearest_kernel(const float* __restrict__ x, float* __restrict__ out, long long planes, int h,
                                        int w, int sf) {
  const int H = h * sf, W = w * sf;
  const long long total = planes * H * W;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int X = int(i % W), Y = int((i / W) % H);
    const long long pl = i / (static_cast<long long>(W) * H);
    out[i] = __ldg(x + (pl * h + Y / sf) * w + X / sf);
  }
}


// ===========================================================================
// Backward pass of the super-resolution network's small layers
// ===========================================================================

// SFT backward (autograd of a = lrelu(x * mul + add), networks/AttResUNet.py:54-58), after the producing dgrad
// has already applied lrelu'(a):  g = dL/d(x * mul + add), NHWC DT [n][npix][ld]
//   gx = g * mul[n][c] (+ resid),  dmul[n][c] += sum_pix g * x,  dadd[n][c] += sum_pix g
// HBM-bound (3-4 tensors streamed once): every thread moves 16-byte vectors of one channel group and keeps the
// group's partial sums in registers; lanes are reduced through shared memory, one atomicAdd per (block, channel).
// blockIdx.y = sample, blockIdx.x = pixel chunk.
template <typename DT>
struct SftVec;
template <>
struct SftVec<__nv_bfloat16> {
  static constexpr int kN = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* v) {
    unpack(*reinterpret_cast<const uint4*>(p), v);
  }
  static __device__ __forceinline__ void unpack(const uint4& r, float* v) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x, v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float* v) {
    uint4 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = r;
  }
};
template <>
struct SftVec<float> {
  static constexpr int kN = 4;
  static __device__ __forceinline__ void load(const float* p, float* v) {
    unpack(*reinterpret_cast<const uint4*>(p), v);
  }
  static __device__ __forceinline__ void unpack(const uint4& r, float* v) {
    v[0] = __uint_as_float(r.x), v[1] = __uint_as_float(r.y), v[2] = __uint_as_float(r.z), v[3] = __uint_as_float(r.w);
  }
  static __device__ __forceinline__ void store(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

// U = pixels a thread has in flight per loop step (all their 16-byte loads are issued before the first use): the kernel
// is bound by bytes in flight per SM (registers cap the resident threads), not by arithmetic.
template <typename DT, int U>
__global__ void __launch_bounds__(256, U == 1 ? 4 : 3)
sft_bwd_kernel(const DT* __restrict__ g, const DT* __restrict__ x, const float* __restrict__ mul,
               const DT* __restrict__ resid, DT* __restrict__ gx, float* __restrict__ dmul, float* __restrict__ dadd,
               int npix, int C, int ld, int pix_per_block, float* __restrict__ slots) {
  constexpr int V = SftVec<DT>::kN;
  extern __shared__ float sm[];            // [lanes][ld] x 2
  const int n = blockIdx.y;
  const int groups = ld / V, lanes = 256 / groups;
  const int cg = threadIdx.x % groups, pl = threadIdx.x / groups;
  const long long base = static_cast<long long>(n) * npix * ld;
  const int p0 = blockIdx.x * pix_per_block, p1 = min(npix, p0 + pix_per_block);
  float m[V], sm_[V], sa_[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = cg * V + i;
    m[i] = c < C ? mul[n * C + c] : 0.f;
    sm_[i] = sa_[i] = 0.f;
  }
  const bool has_resid = resid != nullptr;
  auto ldv = [](const DT* p) { return *reinterpret_cast<const uint4*>(p); };
  auto one = [&](long long i0, const uint4& gr, const uint4& xr, const uint4& rr) {
    float gv[V], xv[V], o[V];
    SftVec<DT>::unpack(gr, gv);
    SftVec<DT>::unpack(xr, xv);
    if (has_resid) {
      SftVec<DT>::unpack(rr, o);
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = fmaf(gv[i], m[i], o[i]);
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = gv[i] * m[i];
    }
#pragma unroll
    for (int i = 0; i < V; ++i) {
      sm_[i] = fmaf(gv[i], xv[i], sm_[i]);
      sa_[i] += gv[i];
    }
    SftVec<DT>::store(gx + i0, o);
  };
  if (pl < lanes) {
    int p = p0 + pl;
    if constexpr (U == 2) {
      for (; p + lanes < p1; p += 2 * lanes) {
        const long long i0 = base + static_cast<long long>(p) * ld + cg * V;
        const long long i1 = i0 + static_cast<long long>(lanes) * ld;
        const uint4 g0 = ldv(g + i0), x0 = ldv(x + i0), g1 = ldv(g + i1), x1 = ldv(x + i1);
        uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
        if (has_resid) r0 = ldv(resid + i0), r1 = ldv(resid + i1);
        one(i0, g0, x0, r0);
        one(i1, g1, x1, r1);
      }
    }
    for (; p < p1; p += lanes) {
      const long long i0 = base + static_cast<long long>(p) * ld + cg * V;
      const uint4 g0 = ldv(g + i0), x0 = ldv(x + i0);
      uint4 r0 = make_uint4(0, 0, 0, 0);
      if (has_resid) r0 = ldv(resid + i0);
      one(i0, g0, x0, r0);
    }
  }
  float* pm = sm;
  float* pa = sm + lanes * ld;
  if (pl < lanes) {
#pragma unroll
    for (int i = 0; i < V; ++i) pm[pl * ld + cg * V + i] = sm_[i], pa[pl * ld + cg * V + i] = sa_[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float tm = 0.f, ta = 0.f;
    for (int l = 0; l < lanes; ++l) tm += pm[l * ld + c], ta += pa[l * ld + c];
    if (slots != nullptr) {
      // deterministic form: slots [pixel chunk][mul | add][n][C], added in chunk order by slot_sum_kernel
      const long long NC = static_cast<long long>(gridDim.y) * C;
      slots[(blockIdx.x * 2ll + 0) * NC + n * C + c] = tm;
      slots[(blockIdx.x * 2ll + 1) * NC + n * C + c] = ta;
    } else {
      atomicAdd(dmul + n * C + c, tm);
      atomicAdd(dadd + n * C + c, ta);
    }
  }
}

// Ordered reduction of per-block / per-sample partial sums (the deterministic forms of the kernels in this file):
//   seg.out[i] += sum_{s < nslots} ws[s * stride + seg.off + i],  s ascending;  blockIdx.y = segment.
struct SlotSegs {
  float* out[4];
  int off[4];
  int count[4];
};
__global__ void __launch_bounds__(256) slot_sum_kernel(const float* __restrict__ ws, int nslots, long long stride, SlotSegs segs) {
  // block = 8 warps x 32 consecutive elements: warp w adds slots w, w + 8, ... in that order (coalesced rows), then the
  // eight warp sums are added in warp order — a fixed summation tree
  __shared__ float part[8][32];
  const int k = blockIdx.y, w = threadIdx.x >> 5, l = threadIdx.x & 31;
  float* out = segs.out[k];
  const float* src = ws + segs.off[k];
  const int count = segs.count[k];
  for (int base = blockIdx.x * 32; base < count; base += gridDim.x * 32) {
    const int i = base + l;
    float a = 0.f;
    if (i < count)
      for (int sl = w; sl < nslots; sl += 8) a += src[sl * stride + i];
    part[w][l] = a;
    __syncthreads();
    if (w == 0 && i < count) {
      float t = part[0][l];
#pragma unroll
      for (int j = 1; j < 8; ++j) t += part[j][l];
      out[i] += t;
    }
    __syncthreads();
  }
}

// AttLayer MLP backward, one block per sample (parameter gradients by atomicAdd over samples):
// recomputes the forward of sft_mlp_kernel, then back-propagates (dmul, dadd) to the four 1x1 convs and to the
// conditioning values.  d_extra[n][e] += dL/d(raw extra) (the sqrt of the masked entries is chained here).
__device__ __forceinline__ void sft_mlp_bwd_body(int n, const float* __restrict__ extra, int E, unsigned sqrt_mask,
                                   const float* __restrict__ w1, const float* __restrict__ b1, int C1,
                                   const float* __restrict__ w2, const float* __restrict__ b2, int C2,
                                   const float* __restrict__ wm, const float* __restrict__ bm,
                                   const float* __restrict__ wa, const float* __restrict__ ba, int C, float alpha,
                                   const float* __restrict__ dmul, const float* __restrict__ dadd, float* __restrict__ gw1,
                                   float* __restrict__ gb1, float* __restrict__ gw2, float* __restrict__ gb2,
                                   float* __restrict__ gwm, float* __restrict__ gbm, float* __restrict__ gwa,
                                   float* __restrict__ gba, float* __restrict__ d_extra, bool plain = false) {
  // plain: every gradient pointer addresses a private slot of this (layer, sample): plain stores instead of atomics
  auto accum = [plain](float* p, float v) {
    if (plain) *p = v;
    else atomicAdd(p, v);
  };
  extern __shared__ float sm[];   // e[E] f1p[C1] f1[C1] f2p[C2] f2[C2] gmp[C] gad[C] gf2[C2] gf1[C1]
  float* e = sm;
  float* f1p = e + E;
  float* f1 = f1p + C1;
  float* f2p = f1 + C1;
  float* f2 = f2p + C2;
  float* gmp = f2 + C2;
  float* gad = gmp + C;
  float* gf2 = gad + C;
  float* gf1 = gf2 + C2;
  const int T = blockDim.x, t = threadIdx.x;
  {
    if (t < E) {
      const float v = extra[n * E + t];
      e[t] = (sqrt_mask & (1u << t)) ? sqrtf(v) : v;
    }
    __syncthreads();
    for (int i = t; i < C1; i += T) {
      float a = b1[i];
      for (int k = 0; k < E; ++k) a = fmaf(w1[i * E + k], e[k], a);
      f1p[i] = a, f1[i] = a > 0.f ? a : a * alpha;
    }
    __syncthreads();
    for (int i = t; i < C2; i += T) {
      float a = b2[i];
      for (int k = 0; k < C1; ++k) a = fmaf(w2[i * C1 + k], f1[k], a);
      f2p[i] = a, f2[i] = a > 0.f ? a : a * alpha;
    }
    __syncthreads();
    for (int c = t; c < C; c += T) {
      float a = bm[c];
      for (int k = 0; k < C2; ++k) a = fmaf(wm[c * C2 + k], f2[k], a);
      const float mu = 1.f / (1.f + expf(-a));
      const float g1 = dmul[n * C + c] * mu * (1.f - mu), g2 = dadd[n * C + c];
      gmp[c] = g1, gad[c] = g2;
      accum(gbm + c, g1), accum(gba + c, g2);
      for (int k = 0; k < C2; ++k) accum(gwm + c * C2 + k, g1 * f2[k]), accum(gwa + c * C2 + k, g2 * f2[k]);
    }
    __syncthreads();
    for (int k = t; k < C2; k += T) {
      float a = 0.f;
      for (int c = 0; c < C; ++c) a = fmaf(wm[c * C2 + k], gmp[c], fmaf(wa[c * C2 + k], gad[c], a));
      a *= f2p[k] > 0.f ? 1.f : alpha;
      gf2[k] = a;
      accum(gb2 + k, a);
      for (int j = 0; j < C1; ++j) accum(gw2 + k * C1 + j, a * f1[j]);
    }
    __syncthreads();
    for (int j = t; j < C1; j += T) {
      float a = 0.f;
      for (int k = 0; k < C2; ++k) a = fmaf(w2[k * C1 + j], gf2[k], a);
      a *= f1p[j] > 0.f ? 1.f : alpha;
      gf1[j] = a;
      accum(gb1 + j, a);
      for (int i = 0; i < E; ++i) accum(gw1 + j * E + i, a * e[i]);
    }
    __syncthreads();
    if (t < E) {
      float a = 0.f;
      for (int j = 0; j < C1; ++j) a = fmaf(w1[j * E + t], gf1[j], a);
      if (sqrt_mask & (1u << t)) a *= 0.5f / fmaxf(e[t], 1e-20f);
      accum(d_extra + n * E + t, a);
    }
  }
}

__global__ void sft_mlp_bwd_kernel(const float* __restrict__ extra, int N, int E, unsigned sqrt_mask,
                                   const float* __restrict__ w1, const float* __restrict__ b1, int C1,
                                   const float* __restrict__ w2, const float* __restrict__ b2, int C2,
                                   const float* __restrict__ wm, const float* __restrict__ bm,
                                   const float* __restrict__ wa, const float* __restrict__ ba, int C, float alpha,
                                   const float* __restrict__ dmul, const float* __restrict__ dadd, float* __restrict__ gw1,
                                   float* __restrict__ gb1, float* __restrict__ gw2, float* __restrict__ gb2,
                                   float* __restrict__ gwm, float* __restrict__ gbm, float* __restrict__ gwa,
                                   float* __restrict__ gba, float* __restrict__ d_extra) {
  sft_mlp_bwd_body(blockIdx.x, extra, E, sqrt_mask, w1, b1, C1, w2, b2, C2, wm, bm, wa, ba, C, alpha, dmul, dadd, gw1, gb1,
                   gw2, gb2, gwm, gbm, gwa, gba, d_extra);
}

__global__ void sft_mlp_bwd_batched_kernel(const SftDesc* __restrict__ descs, const float* __restrict__ extra, int E,
                                           unsigned sqrt_mask, float alpha, float* __restrict__ d_extra) {
  const SftDesc d = descs[blockIdx.y];
  sft_mlp_bwd_body(blockIdx.x, extra, E, sqrt_mask, d.w1, d.b1, d.c1, d.w2, d.b2, d.c2, d.wm, d.bm, d.wa, d.ba, d.c, alpha,
                   d.dmul, d.dadd, d.gw1, d.gb1, d.gw2, d.gb2, d.gwm, d.gbm, d.gwa, d.gba, d_extra);
}

// Deterministic form: block (sample n, layer l) stores its parameter gradients into its own slot
//   ws[n * P + off_l + (w1 | b1 | w2 | b2 | wm | bm | wa | ba)],  P = parameters of all layers, off_l = those before l,
// and its conditioning gradient into ws[N * P + (l * N + n) * E + e]; sft_mlp_slot_reduce_kernel then adds the sample
// slots in sample order into the parameter gradients and the layer slots in layer order into d_extra.
__device__ __forceinline__ long long sft_layer_params(const SftDesc& d, int E) {
  return static_cast<long long>(d.c1) * E + d.c1 + static_cast<long long>(d.c2) * d.c1 + d.c2 +
         2 * (static_cast<long long>(d.c) * d.c2 + d.c);
}

__global__ void sft_mlp_bwd_batched_slots_kernel(const SftDesc* __restrict__ descs, const float* __restrict__ extra, int N,
                                                 int E, unsigned sqrt_mask, float alpha, float* __restrict__ ws,
                                                 long long P) {
  const int n = blockIdx.x, l = blockIdx.y;
  long long off = 0;
  for (int j = 0; j < l; ++j) off += sft_layer_params(descs[j], E);
  const SftDesc d = descs[l];
  float* gw1 = ws + n * P + off;
  float* gb1 = gw1 + d.c1 * E;
  float* gw2 = gb1 + d.c1;
  float* gb2 = gw2 + d.c2 * d.c1;
  float* gwm = gb2 + d.c2;
  float* gbm = gwm + d.c * d.c2;
  float* gwa = gbm + d.c;
  float* gba = gwa + d.c * d.c2;
  float* dslot = ws + static_cast<long long>(N) * P + static_cast<long long>(l) * N * E;
  sft_mlp_bwd_body(n, extra, E, sqrt_mask, d.w1, d.b1, d.c1, d.w2, d.b2, d.c2, d.wm, d.bm, d.wa, d.ba, d.c, alpha, d.dmul,
                   d.dadd, gw1, gb1, gw2, gb2, gwm, gbm, gwa, gba, dslot, true);
}

// blockIdx.y < n_layers: the parameter gradients of that layer; blockIdx.y == n_layers: d_extra.  Block = 8 warps x 32
// consecutive elements; warp w adds slots w, w + 8, ... in order, the warp sums are then added in warp order.
__global__ void __launch_bounds__(256)
sft_mlp_slot_reduce_kernel(const SftDesc* __restrict__ descs, int n_layers, int N, int E, const float* __restrict__ ws,
                           long long P, float* __restrict__ d_extra) {
  __shared__ float part[8][32];
  const int l = blockIdx.y, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool is_extra = l == n_layers;
  long long off = 0, count = static_cast<long long>(N) * E, stride = count;
  const float* src = ws + static_cast<long long>(N) * P;
  int nslots = n_layers;
  SftDesc d{};
  if (!is_extra) {
    for (int j = 0; j < l; ++j) off += sft_layer_params(descs[j], E);
    d = descs[l];
    count = sft_layer_params(d, E), stride = P, src = ws + off, nslots = N;
  }
  for (long long base = blockIdx.x * 32ll; base < count; base += gridDim.x * 32ll) {
    const long long i = base + lane;
    float a = 0.f;
    if (i < count)
      for (int sl = w; sl < nslots; sl += 8) a += src[sl * stride + i];
    part[w][lane] = a;
    __syncthreads();
    if (w == 0 && i < count) {
      float t = part[0][lane];
#pragma unroll
      for (int j = 1; j < 8; ++j) t += part[j][lane];
      float* out = d_extra;
      long long k = i;
      if (!is_extra) {
        // segment of the slot layout w1 | b1 | w2 | b2 | wm | bm | wa | ba
        const long long s1 = static_cast<long long>(d.c1) * E, s2 = s1 + d.c1, s3 = s2 + static_cast<long long>(d.c2) * d.c1,
                        s4 = s3 + d.c2, s5 = s4 + static_cast<long long>(d.c) * d.c2, s6 = s5 + d.c,
                        s7 = s6 + static_cast<long long>(d.c) * d.c2;
        if (i < s1) out = d.gw1;
        else if (i < s2) out = d.gb1, k = i - s1;
        else if (i < s3) out = d.gw2, k = i - s2;
        else if (i < s4) out = d.gb2, k = i - s3;
        else if (i < s5) out = d.gwm, k = i - s4;
        else if (i < s6) out = d.gbm, k = i - s5;
        else if (i < s7) out = d.gwa, k = i - s6;
        else out = d.gba, k = i - s7;
      }
      out[k] += t;
    }
    __syncthreads();
  }
}

// CALayer + skip backward (autograd of out = f * s(mean f) + skip), one CTA per sample:
//   d_f = g * s + dy / npix,  the skip gradient is g itself;  parameter gradients by atomicAdd over samples.
template <typename DT>
__global__ void __launch_bounds__(1024) ca_layer_bwd_kernel(const DT* __restrict__ g, const DT* __restrict__ f, const float* __restrict__ w1,
                                    const float* __restrict__ b1, const float* __restrict__ w2,
                                    const float* __restrict__ b2, DT* __restrict__ df, float* __restrict__ gw1,
                                    float* __restrict__ gb1, float* __restrict__ gw2, float* __restrict__ gb2, int npix,
                                    int C, int R, int ld, float alpha, float* __restrict__ slots) {
  extern __shared__ float sm[];   // part_y[L][C] part_d[L][C] y[C] ds[C] s[C] gsp[C] dy[C] zp[R] z[R] gzp[R]
  const int n = blockIdx.x;
  // deterministic form: sample n stores its parameter gradients into its own slot [w1 | b1 | w2 | b2] (plain stores),
  // slot_sum_kernel adds the slots in sample order
  const bool plain = slots != nullptr;
  if (plain) {
    float* sl = slots + static_cast<long long>(n) * (2 * R * C + R + C);
    gw1 = sl, gb1 = gw1 + R * C, gw2 = gb1 + R, gb2 = gw2 + C * R;
  }
  auto accum = [plain](float* p, float v) {
    if (plain) *p = v;
    else atomicAdd(p, v);
  };
  const long long base = static_cast<long long>(n) * npix * ld;
  const int lanes = blockDim.x / C;
  const int c = threadIdx.x % C, pl = threadIdx.x / C;
  float* part_y = sm;
  float* part_d = part_y + lanes * C;
  float* y = part_d + lanes * C;
  float* ds = y + C;
  float* s = ds + C;
  float* gsp = s + C;
  float* dy = gsp + C;
  float* zp = dy + C;
  float* z = zp + R;
  float* gzp = z + R;
  float py = 0.f, pd = 0.f;
  if (pl < lanes)
    for (int p = pl; p < npix; p += lanes) {
      const float fv = ld_f<DT>(f + base + static_cast<long long>(p) * ld + c);
      py += fv;
      pd += fv * ld_f<DT>(g + base + static_cast<long long>(p) * ld + c);
    }
  if (pl < lanes) part_y[pl * C + c] = py, part_d[pl * C + c] = pd;
  __syncthreads();
  if (threadIdx.x < C) {
    float a = 0.f, b = 0.f;
    for (int l = 0; l < lanes; ++l) a += part_y[l * C + threadIdx.x], b += part_d[l * C + threadIdx.x];
    y[threadIdx.x] = a / float(npix), ds[threadIdx.x] = b;
  }
  __syncthreads();
  if (threadIdx.x < R) {
    float a = b1[threadIdx.x];
    for (int k = 0; k < C; ++k) a = fmaf(w1[threadIdx.x * C + k], y[k], a);
    zp[threadIdx.x] = a, z[threadIdx.x] = a > 0.f ? a : a * alpha;
  }
  __syncthreads();
  if (threadIdx.x < C) {
    float a = b2[threadIdx.x];
    for (int k = 0; k < R; ++k) a = fmaf(w2[threadIdx.x * R + k], z[k], a);
    const float sv = 1.f / (1.f + expf(-a));
    s[threadIdx.x] = sv;
    const float gs = ds[threadIdx.x] * sv * (1.f - sv);
    gsp[threadIdx.x] = gs;
    accum(gb2 + threadIdx.x, gs);
    for (int k = 0; k < R; ++k) accum(gw2 + threadIdx.x * R + k, gs * z[k]);
  }
  __syncthreads();
  if (threadIdx.x < R) {
    float a = 0.f;
    for (int cc = 0; cc < C; ++cc) a = fmaf(w2[cc * R + threadIdx.x], gsp[cc], a);
    a *= zp[threadIdx.x] > 0.f ? 1.f : alpha;
    gzp[threadIdx.x] = a;
    accum(gb1 + threadIdx.x, a);
    for (int cc = 0; cc < C; ++cc) accum(gw1 + threadIdx.x * C + cc, a * y[cc]);
  }
  __syncthreads();
  if (threadIdx.x < C) {
    float a = 0.f;
    for (int k = 0; k < R; ++k) a = fmaf(w1[k * C + threadIdx.x], gzp[k], a);
    dy[threadIdx.x] = a / float(npix);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < npix * ld; i += blockDim.x) {
    const int cc = i % ld;
    float v = 0.f;
    if (cc < C) v = ld_f<DT>(g + base + i) * s[cc] + dy[cc];
    st_f<DT>(df + base + i, v);
  }
}

// Backward of gap_head_kernel: the gradient w.r.t. every pixel of plane (n, c) is
//   g[n][c] * head'(.) / hw,  head' = out (exp head, inside the clamp range) or 1 - out^2 (tanh); NHWC DT output.
template <typename DT>
__global__ void gap_head_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ outv, int C, int hw,
                                    unsigned exp_mask, unsigned tanh_mask, float out_lo, float out_hi,
                                    DT* __restrict__ gx, int ld, long long total) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = int(i % ld);
    const long long n = i / (static_cast<long long>(ld) * hw);
    float v = 0.f;
    if (c < C) {
      const float o = outv[n * C + c];
      float d = 1.f;
      if (exp_mask & (1u << c)) d = (o > out_lo && o < out_hi) ? o : 0.f;
      if (tanh_mask & (1u << c)) d = 1.f - o * o;
      v = gout[n * C + c] * d / float(hw);
    }
    st_f<DT>(gx + i, v);
  }
}

// Weight gradient of the KNet head (9x9, stride 4, pad 4): one thread per (weight element, sample), atomicAdd over samples;
// deterministic form: the thread stores its sum into the sample's slot, slot_sum_kernel adds the slots in sample order.
template <typename DT>
__global__ void knet_head_wgrad_kernel(const float* __restrict__ x, const DT* __restrict__ g, float* __restrict__ gw,
                                       int N, int C, int H, int W, int OH, int OW, int cout, int ld,
                                       float* __restrict__ slots) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cout * C * 81) return;
  const int s = i % 9, r = (i / 9) % 9, c = (i / 81) % C, co = i / (81 * C);
  float acc = 0.f;
  const int n = blockIdx.y;
  {
    for (int oy = 0; oy < OH; ++oy) {
      const int iy = oy * 4 - 4 + r;
      if (iy < 0 || iy >= H) continue;
      for (int ox = 0; ox < OW; ++ox) {
        const int ix = ox * 4 - 4 + s;
        if (ix < 0 || ix >= W) continue;
        acc = fmaf(ld_f<DT>(g + ((static_cast<long long>(n) * OH + oy) * OW + ox) * ld + co),
                   __ldg(x + ((static_cast<long long>(n) * C + c) * H + iy) * W + ix), acc);
      }
    }
  }
  if (slots != nullptr) slots[static_cast<long long>(n) * (cout * C * 81) + i] = acc;
  else atomicAdd(gw + i, acc);
}

}  // namespace vk

using namespace vk;

// CALayer kernels run one CTA per sample: the SM has nothing else to do, so use up to 1024 threads (a multiple of the
// channel count) as soon as every pixel lane has a few pixels to sum
static int ca_threads(int npix, int c) {
  const int cap = static_cast<long long>(npix) * c >= 8192 ? 1024 : 256;
  return std::max(c, cap / c * c);
}

#define VK_ST(s) reinterpret_cast<cudaStream_t>(s)
#define VK_LAUNCHED()                                        \
  g_launch_count.fetch_add(1, std::memory_order_relaxed);    \
  return int(cudaGetLastError())

extern "C" int vk_knet_head(int32_t dtype, const float* x, const float* w, void* out, int32_t n, int32_t c, int32_t h,
                            int32_t wd, int32_t cout, int32_t ld, void* stream) {
  if (!x || !w || !out || n <= 0 || c <= 0 || h <= 0 || wd <= 0 || cout <= 0 || cout > ld) return VK_E_BADARG;
  const int oh = (h - 1) / 4 + 1, ow = (wd - 1) / 4 + 1;
  const size_t smem = size_t(81) * ld * sizeof(float);               // one input channel's filter slice, transposed
  if (ld > 256 || smem > 48 * 1024) return VK_E_BADARG;
  const long long per_block = static_cast<long long>(256 / ld) * 4;  // pixels per block step (4 groups per staged slice)
  const long long npix = static_cast<long long>(n) * oh * ow;
  const int grid = int(std::min<long long>((npix + per_block - 1) / per_block, 148 * 8));
  if (dtype == VK_BF16)
    knet_head_kernel<__nv_bfloat16><<<grid, 256, smem, VK_ST(stream)>>>(x, w, reinterpret_cast<__nv_bfloat16*>(out), n, c,
                                                                       h, wd, oh, ow, cout, ld);
  else if (dtype == VK_TF32)
    knet_head_kernel<float><<<grid, 256, smem, VK_ST(stream)>>>(x, w, reinterpret_cast<float*>(out), n, c, h, wd, oh, ow,
                                                               cout, ld);
  else
    return VK_E_BADARG;
  VK_LAUNCHED();
}

extern "C" int vk_ca_layer(int32_t dtype, const void* f, const void* skip, const float* w1, const float* b1,
                           const float* w2, const float* b2, void* out, int32_t n, int32_t npix, int32_t c, int32_t r,
                           int32_t ld, float alpha, void* stream) {
  if (!f || !skip || !w1 || !b1 || !w2 || !b2 || !out || n <= 0 || npix <= 0 || c <= 0 || r <= 0 || c > ld || c > 256)
    return VK_E_BADARG;
  const int threads = ca_threads(npix, c);
  const size_t smem = (size_t(threads / c) * c + 2 * c + r) * sizeof(float);
  if (dtype == VK_BF16)
    ca_layer_kernel<__nv_bfloat16><<<n, threads, smem, VK_ST(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(f), reinterpret_cast<const __nv_bfloat16*>(skip), w1, b1, w2, b2,
        reinterpret_cast<__nv_bfloat16*>(out), npix, c, r, ld, alpha);
  else if (dtype == VK_TF32)
    ca_layer_kernel<float><<<n, threads, smem, VK_ST(stream)>>>(reinterpret_cast<const float*>(f),
                                                              reinterpret_cast<const float*>(skip), w1, b1, w2, b2,
                                                              reinterpret_cast<float*>(out), npix, c, r, ld, alpha);
  else
    return VK_E_BADARG;
  VK_LAUNCHED();
}

extern "C" int vk_gap_head(const float* x, int32_t n, int32_t c, int32_t hw, uint32_t exp_mask, uint32_t tanh_mask,
                           float lo, float hi, float* out, void* stream) {
  if (!x || !out || n <= 0 || c <= 0 || c > 32 || hw <= 0) return VK_E_BADARG;
  gap_head_kernel<<<n * c, 256, 0, VK_ST(stream)>>>(x, hw, c, exp_mask, tanh_mask, lo, hi, out);
  VK_LAUNCHED();
}

extern "C" int vk_sft_mlp(const float* extra, int32_t n, int32_t e, uint32_t sqrt_mask, const float* w1,
                          const float* b1, int32_t c1, const float* w2, const float* b2, int32_t c2, const float* wm,
                          const float* bm, const float* wa, const float* ba, int32_t c, float alpha, float* mul,
                          float* add, void* stream) {
  if (!extra || !w1 || !b1 || !w2 || !b2 || !wm || !bm || !wa || !ba || !mul || !add) return VK_E_BADARG;
  if (n <= 0 || e <= 0 || e > 32 || c1 <= 0 || c2 <= 0 || c <= 0) return VK_E_BADARG;
  const size_t smem = size_t(e + c1 + c2) * sizeof(float);
  sft_mlp_kernel<<<n, 128, smem, VK_ST(stream)>>>(extra, e, sqrt_mask, w1, b1, c1, w2, b2, c2, wm, bm, wa, ba, c, alpha,
                                                 mul, add);
  VK_LAUNCHED();
}

extern "C" int vk_sft_mlp_batched(const void* descs_dev, int32_t n_layers, int32_t max_c, const float* extra, int32_t n,
                                  int32_t e, uint32_t sqrt_mask, float alpha, void* stream) {
  if (!descs_dev || !extra || n_layers <= 0 || n <= 0 || e <= 0 || e > 32 || max_c <= 0) return VK_E_BADARG;
  const size_t smem = size_t(e + max_c) * sizeof(float);               // c1 + c2 <= c / 8 + c / 4 < max_c
  sft_mlp_batched_kernel<<<dim3(n, n_layers), 128, smem, VK_ST(stream)>>>(reinterpret_cast<const SftDesc*>(descs_dev),
                                                                         extra, e, sqrt_mask, alpha);
  VK_LAUNCHED();
}

extern "C" int vk_sft_mlp_bwd_batched(const void* descs_dev, int32_t n_layers, int32_t max_c, const float* extra,
                                      int32_t n, int32_t e, uint32_t sqrt_mask, float alpha, float* d_extra,
                                      void* stream) {
  if (!descs_dev || !extra || !d_extra || n_layers <= 0 || n <= 0 || e <= 0 || e > 32 || max_c <= 0) return VK_E_BADARG;
  const size_t smem = size_t(e + 5 * max_c) * sizeof(float);           // 3 c1 + 3 c2 + 2 c <= 5 c
  sft_mlp_bwd_batched_kernel<<<dim3(n, n_layers), 128, smem, VK_ST(stream)>>>(
      reinterpret_cast<const SftDesc*>(descs_dev), extra, e, sqrt_mask, alpha, d_extra);
  VK_LAUNCHED();
}

extern "C" int vk_sft_mlp_bwd_batched_det(const void* descs_dev, int32_t n_layers, int32_t max_c, const float* extra,
                                          int32_t n, int32_t e, uint32_t sqrt_mask, float alpha, float* d_extra,
                                          int64_t params_per_sample, float* ws, int64_t ws_floats, void* stream) {
  if (!descs_dev || !extra || !d_extra || !ws || n_layers <= 0 || n <= 0 || e <= 0 || e > 32 || max_c <= 0 ||
      params_per_sample <= 0)
    return VK_E_BADARG;
  if (ws_floats < static_cast<int64_t>(n) * params_per_sample + static_cast<int64_t>(n_layers) * n * e) return VK_E_BADARG;
  const size_t smem = size_t(e + 5 * max_c) * sizeof(float);
  const SftDesc* descs = reinterpret_cast<const SftDesc*>(descs_dev);
  sft_mlp_bwd_batched_slots_kernel<<<dim3(n, n_layers), 128, smem, VK_ST(stream)>>>(descs, extra, n, e, sqrt_mask, alpha, ws,
                                                                                  params_per_sample);
  // the largest layer has fewer than 3 max_c^2 / 4 parameters: 64 blocks of 32 elements per grid step are plenty
  sft_mlp_slot_reduce_kernel<<<dim3(64, n_layers + 1), 256, 0, VK_ST(stream)>>>(descs, n_layers, n, e, ws, params_per_sample,
                                                                              d_extra);
  g_launch_count.fetch_add(2, std::memory_order_relaxed);
  return int(cudaGetLastError());
}

extern "C" uint32_t vk_sizeof_sft_desc(void) { return uint32_t(sizeof(SftDesc)); }

extern "C" int vk_upsample_nearest(const float* x, float* out, int32_t n, int32_t c, int32_t h, int32_t w, int32_t sf,
                                   void* stream) {
  if (!x || !out || n <= 0 || c <= 0 || h <= 0 || w <= 0 || sf <= 0) return VK_E_BADARG;
  const long long total = static_cast<long long>(n) * c * h * sf * w * sf;
  const int grid = int(std::min<long long>((total + 255) / 256, 148 * 8));
  upsample_nearest_kernel<<<grid, 256, 0, VK_ST(stream)>>>(x, out, static_cast<long long>(n) * c, h, w, sf);
  VK_LAUNCHED();
}

namespace {
constexpr int kSftBwdPixPerBlock = 512;

// out[k][i] += the `nslots` partials of segment k, in slot order (second launch of the deterministic forms)
int launch_slot_sum(const float* ws, int nslots, long long stride, const SlotSegs& segs, int nseg, cudaStream_t st) {
  int most = 0;
  for (int k = 0; k < nseg; ++k) most = std::max(most, segs.count[k]);
  slot_sum_kernel<<<dim3((most + 31) / 32, nseg), 256, 0, st>>>(ws, nslots, stride, segs);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return int(cudaGetLastError());
}

int sft_bwd_launch(int32_t dtype, const void* g, const void* x, const float* mul, const void* resid, void* gx,
                   float* dmul, float* dadd, int32_t n, int32_t npix, int32_t c, int32_t ld, bool det, float* ws,
                   int64_t ws_floats, void* stream) {
  if (!g || !x || !mul || !gx || !dmul || !dadd || n <= 0 || npix <= 0 || c <= 0 || c > ld) return VK_E_BADARG;
  const int vec = dtype == VK_BF16 ? 8 : 4;
  if (ld % vec != 0 || ld / vec > 256) return VK_E_BADARG;
  const int ppb = kSftBwdPixPerBlock;
  const size_t smem = size_t(2) * (256 / (ld / vec)) * ld * sizeof(float);
  if (smem > 48 * 1024) return VK_E_BADARG;
  dim3 grid((npix + ppb - 1) / ppb, n);
  const long long nc = static_cast<long long>(n) * c;
  if (det && (!ws || ws_floats < static_cast<long long>(grid.x) * 2 * nc)) return VK_E_BADARG;
  float* slots = det ? ws : nullptr;
  // pixels in flight per thread: 2 for bf16, 1 for fp32 storage (measured, profiles/r02_sft_bwd_unroll.txt);
  // VK_SFT_BWD_UNROLL=1|2 overrides (tuning knob)
  static const int forced = [] {
    const char* e = std::getenv("VK_SFT_BWD_UNROLL");
    return e && (e[0] == '1' || e[0] == '2') ? e[0] - '0' : 0;
  }();
  const int unroll = forced ? forced : (dtype == VK_BF16 ? 2 : 1);
  auto launch = [&](auto tag, auto u) {
    using DT = decltype(tag);
    sft_bwd_kernel<DT, decltype(u)::value><<<grid, 256, smem, VK_ST(stream)>>>(
        reinterpret_cast<const DT*>(g), reinterpret_cast<const DT*>(x), mul, reinterpret_cast<const DT*>(resid),
        reinterpret_cast<DT*>(gx), dmul, dadd, npix, c, ld, ppb, slots);
  };
  using U1 = std::integral_constant<int, 1>;
  using U2 = std::integral_constant<int, 2>;
  if (dtype == VK_BF16)
    unroll == 1 ? launch(__nv_bfloat16{}, U1{}) : launch(__nv_bfloat16{}, U2{});
  else if (dtype == VK_TF32)
    unroll == 1 ? launch(float{}, U1{}) : launch(float{}, U2{});
  else
    return VK_E_BADARG;
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  if (!det) return int(cudaGetLastError());
  SlotSegs segs{};
  segs.out[0] = dmul, segs.off[0] = 0, segs.count[0] = int(nc);
  segs.out[1] = dadd, segs.off[1] = int(nc), segs.count[1] = int(nc);
  return launch_slot_sum(ws, int(grid.x), 2 * nc, segs, 2, VK_ST(stream));
}
}  // namespace

extern "C" int vk_sft_bwd(int32_t dtype, const void* g, const void* x, const float* mul, const void* resid, void* gx,
                          float* dmul, float* dadd, int32_t n, int32_t npix, int32_t c, int32_t ld, void* stream) {
  return sft_bwd_launch(dtype, g, x, mul, resid, gx, dmul, dadd, n, npix, c, ld, false, nullptr, 0, stream);
}

extern "C" int64_t vk_sft_bwd_det_ws_floats(int32_t n, int32_t npix, int32_t c) {
  if (n <= 0 || npix <= 0 || c <= 0) return -1;
  return static_cast<int64_t>((npix + kSftBwdPixPerBlock - 1) / kSftBwdPixPerBlock) * 2 * n * c;
}

extern "C" int vk_sft_bwd_det(int32_t dtype, const void* g, const void* x, const float* mul, const void* resid, void* gx,
                              float* dmul, float* dadd, int32_t n, int32_t npix, int32_t c, int32_t ld, float* ws,
                              int64_t ws_floats, void* stream) {
  return sft_bwd_launch(dtype, g, x, mul, resid, gx, dmul, dadd, n, npix, c, ld, true, ws, ws_floats, stream);
}

extern "C" int vk_sft_mlp_bwd(const float* extra, int32_t n, int32_t e, uint32_t sqrt_mask, const float* w1,
                              const float* b1, int32_t c1, const float* w2, const float* b2, int32_t c2,
                              const float* wm, const float* bm, const float* wa, const float* ba, int32_t c, float alpha,
                              const float* dmul, const float* dadd, float* gw1, float* gb1, float* gw2, float* gb2,
                              float* gwm, float* gbm, float* gwa, float* gba, float* d_extra, void* stream) {
  if (!extra || !w1 || !b1 || !w2 || !b2 || !wm || !bm || !wa || !ba || !dmul || !dadd || !gw1 || !gb1 || !gw2 ||
      !gb2 || !gwm || !gbm || !gwa || !gba || !d_extra)
    return VK_E_BADARG;
  if (n <= 0 || e <= 0 || e > 32 || c1 <= 0 || c2 <= 0 || c <= 0) return VK_E_BADARG;
  const size_t smem = size_t(e + 3 * c1 + 3 * c2 + 2 * c) * sizeof(float);
  sft_mlp_bwd_kernel<<<n, 128, smem, VK_ST(stream)>>>(extra, n, e, sqrt_mask, w1, b1, c1, w2, b2, c2, wm, bm, wa, ba, c,
                                                     alpha, dmul, dadd, gw1, gb1, gw2, gb2, gwm, gbm, gwa, gba, d_extra);
  VK_LAUNCHED();
}

namespace {
int ca_layer_bwd_launch(int32_t dtype, const void* g, const void* f, const float* w1, const float* b1, const float* w2,
                        const float* b2, void* df, float* gw1, float* gb1, float* gw2, float* gb2, int32_t n, int32_t npix,
                        int32_t c, int32_t r, int32_t ld, float alpha, bool det, float* ws, int64_t ws_floats,
                        void* stream) {
  if (!g || !f || !w1 || !b1 || !w2 || !b2 || !df || !gw1 || !gb1 || !gw2 || !gb2) return VK_E_BADARG;
  if (n <= 0 || npix <= 0 || c <= 0 || r <= 0 || c > ld || c > 256) return VK_E_BADARG;
  const int slot = 2 * r * c + r + c;
  if (det && (!ws || ws_floats < static_cast<int64_t>(n) * slot)) return VK_E_BADARG;
  float* slots = det ? ws : nullptr;
  const int threads = ca_threads(npix, c);
  const size_t smem = (size_t(2 * (threads / c)) * c + 5 * c + 3 * r) * sizeof(float);
  if (dtype == VK_BF16)
    ca_layer_bwd_kernel<__nv_bfloat16><<<n, threads, smem, VK_ST(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(g), reinterpret_cast<const __nv_bfloat16*>(f), w1, b1, w2, b2,
        reinterpret_cast<__nv_bfloat16*>(df), gw1, gb1, gw2, gb2, npix, c, r, ld, alpha, slots);
  else if (dtype == VK_TF32)
    ca_layer_bwd_kernel<float><<<n, threads, smem, VK_ST(stream)>>>(reinterpret_cast<const float*>(g),
                                                                  reinterpret_cast<const float*>(f), w1, b1, w2, b2,
                                                                  reinterpret_cast<float*>(df), gw1, gb1, gw2, gb2,
                                                                  npix, c, r, ld, alpha, slots);
  else
    return VK_E_BADARG;
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  if (!det) return int(cudaGetLastError());
  SlotSegs segs{};                     // slot layout of ca_layer_bwd_kernel: [w1 | b1 | w2 | b2]
  segs.out[0] = gw1, segs.off[0] = 0, segs.count[0] = r * c;
  segs.out[1] = gb1, segs.off[1] = r * c, segs.count[1] = r;
  segs.out[2] = gw2, segs.off[2] = r * c + r, segs.count[2] = c * r;
  segs.out[3] = gb2, segs.off[3] = 2 * r * c + r, segs.count[3] = c;
  return launch_slot_sum(ws, n, slot, segs, 4, VK_ST(stream));
}
}  // namespace

extern "C" int vk_ca_layer_bwd(int32_t dtype, const void* g, const void* f, const float* w1, const float* b1,
                               const float* w2, const float* b2, void* df, float* gw1, float* gb1, float* gw2,
                               float* gb2, int32_t n, int32_t npix, int32_t c, int32_t r, int32_t ld, float alpha,
                               void* stream) {
  return ca_layer_bwd_launch(dtype, g, f, w1, b1, w2, b2, df, gw1, gb1, gw2, gb2, n, npix, c, r, ld, alpha, false, nullptr,
                             0, stream);
}

extern "C" int vk_ca_layer_bwd_det(int32_t dtype, const void* g, const void* f, const float* w1, const float* b1,
                                   const float* w2, const float* b2, void* df, float* gw1, float* gb1, float* gw2,
                                   float* gb2, int32_t n, int32_t npix, int32_t c, int32_t r, int32_t ld, float alpha,
                                   float* ws, int64_t ws_floats, void* stream) {
  return ca_layer_bwd_launch(dtype, g, f, w1, b1, w2, b2, df, gw1, gb1, gw2, gb2, n, npix, c, r, ld, alpha, true, ws,
                             ws_floats, stream);
}

extern "C" int vk_gap_head_bwd(int32_t dtype, const float* gout, const float* outv, int32_t n, int32_t c, int32_t hw,
                               uint32_t exp_mask, uint32_t tanh_mask, float lo, float hi, void* gx, int32_t ld,
                               void* stream) {
  if (!gout || !outv || !gx || n <= 0 || c <= 0 || c > 32 || hw <= 0 || c > ld) return VK_E_BADARG;
  const long long total = static_cast<long long>(n) * hw * ld;
  const int grid = int(std::min<long long>((total + 255) / 256, 148 * 8));
  const float olo = expf(lo), ohi = expf(hi);
  if (dtype == VK_BF16)
    gap_head_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, VK_ST(stream)>>>(gout, outv, c, hw, exp_mask, tanh_mask, olo, ohi,
                                                                       reinterpret_cast<__nv_bfloat16*>(gx), ld, total);
  else if (dtype == VK_TF32)
    gap_head_bwd_kernel<float><<<grid, 256, 0, VK_ST(stream)>>>(gout, outv, c, hw, exp_mask, tanh_mask, olo, ohi,
                                                               reinterpret_cast<float*>(gx), ld, total);
  else
    return VK_E_BADARG;
  VK_LAUNCHED();
}

namespace {
int knet_head_wgrad_launch(int32_t dtype, const float* x, const void* g, float* gw, int32_t n, int32_t c, int32_t h,
                           int32_t wd, int32_t cout, int32_t ld, bool det, float* ws, int64_t ws_floats, void* stream) {
  if (!x || !g || !gw || n <= 0 || c <= 0 || h <= 0 || wd <= 0 || cout <= 0 || cout > ld) return VK_E_BADARG;
  const int oh = (h - 1) / 4 + 1, ow = (wd - 1) / 4 + 1;
  const int total = cout * c * 81;
  if (det && (!ws || ws_floats < static_cast<int64_t>(n) * total)) return VK_E_BADARG;
  float* slots = det ? ws : nullptr;
  const dim3 grid((total + 127) / 128, n);
  if (dtype == VK_BF16)
    knet_head_wgrad_kernel<__nv_bfloat16><<<grid, 128, 0, VK_ST(stream)>>>(
        x, reinterpret_cast<const __nv_bfloat16*>(g), gw, n, c, h, wd, oh, ow, cout, ld, slots);
  else if (dtype == VK_TF32)
    knet_head_wgrad_kernel<float><<<grid, 128, 0, VK_ST(stream)>>>(x, reinterpret_cast<const float*>(g), gw, n, c, h, wd, oh,
                                                                 ow, cout, ld, slots);
  else
    return VK_E_BADARG;
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  if (!det) return int(cudaGetLastError());
  SlotSegs segs{};
  segs.out[0] = gw, segs.off[0] = 0, segs.count[0] = total;
  return launch_slot_sum(ws, n, total, segs, 1, VK_ST(stream));
}
}  // namespace

extern "C" int vk_knet_head_wgrad(int32_t dtype, const float* x, const void* g, float* gw, int32_t n, int32_t c,
                                  int32_t h, int32_t wd, int32_t cout, int32_t ld, void* stream) {
  return knet_head_wgrad_launch(dtype, x, g, gw, n, c, h, wd, cout, ld, false, nullptr, 0, stream);
}

extern "C" int vk_knet_head_wgrad_det(int32_t dtype, const float* x, const void* g, float* gw, int32_t n, int32_t c,
                                      int32_t h, int32_t wd, int32_t cout, int32_t ld, float* ws, int64_t ws_floats,
                                      void* stream) {
  return knet_head_wgrad_launch(dtype, x, g, gw, n, c, h, wd, cout, ld, true, ws, ws_floats, stream);
}
