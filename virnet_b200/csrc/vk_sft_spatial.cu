// vk_sft_spatial.cu — SFT modulation with SPATIALLY VARYING conditioning maps (sm_100a).
//
// The AttLayer of the reference (networks/AttResUNet.py:11-32) is a per-pixel MLP on the conditioning maps,
//   f1 = lrelu(W1 e + b1), f2 = lrelu(W2 f1 + b2), mul = sigmoid(Wm f2 + bm), add = Wa f2 + ba,
// and AttResBlock modulates its features with it before each conv: a = lrelu(x * mul + add) (:54-58).  When the maps
// are per-sample constants (the shipped SISR config: kinfo.repeat + GAP'd sigma, networks/VIRNet.py:89-92) the MLP
// runs once per sample (vk_sft_mlp) and the modulation is a conv epilogue.  This file serves the other ctor-legal
// configurations: `noise_avg=False` (JPEG-noise SISR, networks/VIRNet.py:93-95: a per-pixel sigma map next to the
// constant kernel code) and VIRAttResUNet with extra_mode 'Down' / 'Both' (per-pixel sigma map).
//
// Conditioning value of channel e at pixel (y, x) of an h x w feature grid (AttResUNet.py:147-168):
//   full-res padded coordinate  Y = floor(y * Hp / h)         (F.interpolate(..., size, mode='nearest'))
//   reflect back into the image ry = Y < Hh ? Y : 2 (Hh-1) - Y (util_net.pad_input, bottom / right only)
//   e <  ec : cst[n][e]                                        (kinfo.repeat)
//   e >= ec : map[n][e-ec][ry / esf][rx / esf]                 (nearest x sf of the LR sigma map, or esf = 1)
//   sqrt applied to the channels in sqrt_mask                  (VIRNet.py:44,92,94)
//
// One thread per pixel; the AttLayer weights live in shared memory (broadcast reads), f1 / f2 in registers.
#include <algorithm>
#include <cstdio>

#include "../../include/virnet_b200.h"
#include "vk_common.cuh"
#include "vk_host.h"

namespace vk {

namespace {
constexpr int kMaxE = 8;

struct ExtraSrc {
  const float* cst;   // [n][ec] or null
  const float* map;   // [n][em][eh][ew] or null
  int ec, em, eh, ew, esf;
  unsigned sqrt_mask;
  int hh, ww;         // un-padded full-resolution size
  int hp, wp;         // padded full-resolution size
};

__device__ __forceinline__ int reflect_far(int i, int n) { return i < n ? i : 2 * (n - 1) - i; }

// conditioning values of one pixel; (Y, X) on the padded full-resolution grid
__device__ __forceinline__ void load_extra(const ExtraSrc& s, int n, int Y, int X, float (&ex)[kMaxE]) {
  const int ey = reflect_far(Y, s.hh) / s.esf, exx = reflect_far(X, s.ww) / s.esf;
#pragma unroll
  for (int e = 0; e < kMaxE; ++e) {
    float v = 0.f;
    if (e < s.ec) {
      v = __ldg(s.cst + static_cast<long long>(n) * s.ec + e);
    } else if (e < s.ec + s.em) {
      v = __ldg(s.map + ((static_cast<long long>(n) * s.em + (e - s.ec)) * s.eh + ey) * s.ew + exx);
    }
    if (s.sqrt_mask & (1u << e)) v = sqrtf(v);
    ex[e] = v;
  }
}

template <typename DT>
__device__ __forceinline__ void unpack16(const uint4& raw, float (&f)[16 / sizeof(DT)]) {
  if constexpr (sizeof(DT) == 4) {
    f[0] = __uint_as_float(raw.x), f[1] = __uint_as_float(raw.y), f[2] = __uint_as_float(raw.z), f[3] = __uint_as_float(raw.w);
  } else {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) f[2 * i] = __uint_as_float(w[i] << 16), f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}
template <typename DT>
__device__ __forceinline__ uint4 pack16(const float (&f)[16 / sizeof(DT)]) {
  if constexpr (sizeof(DT) == 4) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
  } else {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h2);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
}
__device__ __forceinline__ float lrelu_f(float v, float a) { return v > 0.f ? v : v * a; }
__device__ __forceinline__ float round_tf32_f(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}

struct SftW {
  const float *w1, *b1, *w2, *b2, *wm, *bm, *wa, *ba;
  int c1, c2, c;
};

// shared-memory image of one AttLayer: w1 [c1][E] | b1 | w2 [c2][c1] | b2 | wm [c][C2P] | bm | wa [c][C2P] | ba
// (wm / wa rows padded to C2P so the inner loops are compile-time sized)
template <int C2P>
struct SftSmem {
  static constexpr int C1P = C2P / 2;
  __device__ static int floats(int c, int E) { return C1P * kMaxE + C1P + C2P * C1P + C2P + 2 * (c * C2P + c); }
};

template <int C2P>
__device__ __forceinline__ void stage_weights(float* s, const SftW& w, int E) {
  constexpr int C1P = C2P / 2;
  float* w1 = s;
  float* b1 = w1 + C1P * kMaxE;
  float* w2 = b1 + C1P;
  float* b2 = w2 + C2P * C1P;
  float* wm = b2 + C2P;
  float* bm = wm + w.c * C2P;
  float* wa = bm + w.c;
  float* ba = wa + w.c * C2P;
  for (int i = threadIdx.x; i < C1P * kMaxE; i += blockDim.x) {
    const int r = i / kMaxE, e = i % kMaxE;
    w1[i] = (r < w.c1 && e < E) ? __ldg(w.w1 + r * E + e) : 0.f;
  }
  for (int i = threadIdx.x; i < C1P; i += blockDim.x) b1[i] = i < w.c1 ? __ldg(w.b1 + i) : 0.f;
  for (int i = threadIdx.x; i < C2P * C1P; i += blockDim.x) {
    const int r = i / C1P, k = i % C1P;
    w2[i] = (r < w.c2 && k < w.c1) ? __ldg(w.w2 + r * w.c1 + k) : 0.f;
  }
  for (int i = threadIdx.x; i < C2P; i += blockDim.x) b2[i] = i < w.c2 ? __ldg(w.b2 + i) : 0.f;
  for (int i = threadIdx.x; i < w.c * C2P; i += blockDim.x) {
    const int r = i / C2P, k = i % C2P;
    wm[i] = k < w.c2 ? __ldg(w.wm + r * w.c2 + k) : 0.f;
    wa[i] = k < w.c2 ? __ldg(w.wa + r * w.c2 + k) : 0.f;
  }
  for (int i = threadIdx.x; i < w.c; i += blockDim.x) bm[i] = __ldg(w.bm + i), ba[i] = __ldg(w.ba + i);
}

// f1 = lrelu(W1 e + b1), f2 = lrelu(W2 f1 + b2) of one pixel (padding entries come out as lrelu(0) = 0)
template <int C2P>
__device__ __forceinline__ void mlp_hidden(const float* s, const float (&ex)[kMaxE], float alpha,
                                           float (&f1)[C2P / 2], float (&f2)[C2P]) {
  constexpr int C1P = C2P / 2;
  const float* w1 = s;
  const float* b1 = w1 + C1P * kMaxE;
  const float* w2 = b1 + C1P;
  const float* b2 = w2 + C2P * C1P;
#pragma unroll
  for (int i = 0; i < C1P; ++i) {
    float a = b1[i];
#pragma unroll
    for (int e = 0; e < kMaxE; ++e) a = fmaf(w1[i * kMaxE + e], ex[e], a);
    f1[i] = lrelu_f(a, alpha);
  }
#pragma unroll
  for (int j = 0; j < C2P; ++j) {
    float a = b2[j];
#pragma unroll
    for (int i = 0; i < C1P; ++i) a = fmaf(w2[j * C1P + i], f1[i], a);
    f2[j] = lrelu_f(a, alpha);
  }
}

// out = lrelu(x * mul + add), one thread per pixel
template <typename DT, int C2P>
__global__ void __launch_bounds__(128)
sft_apply_kernel(const DT* __restrict__ x, DT* __restrict__ out, int ld, int N, int h, int w, ExtraSrc src, SftW wt,
                 float alpha, int rnd) {
  extern __shared__ __align__(16) float smem_f[];
  const int E = src.ec + src.em;
  stage_weights<C2P>(smem_f, wt, E);
  __syncthreads();
  constexpr int C1P = C2P / 2;
  constexpr int V = 16 / int(sizeof(DT));
  const float* wm = smem_f + C1P * kMaxE + C1P + C2P * C1P + C2P;
  const float* bm = wm + wt.c * C2P;
  const float* wa = bm + wt.c;
  const float* ba = wa + wt.c * C2P;
  const long long total = static_cast<long long>(N) * h * w;
  const int sy = src.hp / h, sx = src.wp / w;          // nearest: floor(y * Hp / h) with Hp a multiple of h
  for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < total;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int px = int(p % w), py = int((p / w) % h), n = int(p / (static_cast<long long>(w) * h));
    float ex[kMaxE], f1[C1P], f2[C2P];
    load_extra(src, n, py * sy, px * sx, ex);
    mlp_hidden<C2P>(smem_f, ex, alpha, f1, f2);
    const DT* xr = x + p * ld;
    DT* orow = out + p * ld;
    for (int c0 = 0; c0 < ld; c0 += V) {
      float v[V];
      unpack16<DT>(*reinterpret_cast<const uint4*>(xr + c0), v);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int c = c0 + i;
        float r = 0.f;
        if (c < wt.c) {
          float am = bm[c], aa = ba[c];
          const float4* wm4 = reinterpret_cast<const float4*>(wm + c * C2P);
          const float4* wa4 = reinterpret_cast<const float4*>(wa + c * C2P);
#pragma unroll
          for (int j4 = 0; j4 < C2P / 4; ++j4) {
            const float4 m = wm4[j4], a = wa4[j4];
            am = fmaf(m.x, f2[4 * j4], am), am = fmaf(m.y, f2[4 * j4 + 1], am);
            am = fmaf(m.z, f2[4 * j4 + 2], am), am = fmaf(m.w, f2[4 * j4 + 3], am);
            aa = fmaf(a.x, f2[4 * j4], aa), aa = fmaf(a.y, f2[4 * j4 + 1], aa);
            aa = fmaf(a.z, f2[4 * j4 + 2], aa), aa = fmaf(a.w, f2[4 * j4 + 3], aa);
          }
          const float mul = 1.f / (1.f + __expf(-am));
          r = lrelu_f(fmaf(v[i], mul, aa), alpha);
          if (sizeof(DT) == 4 && rnd) r = round_tf32_f(r);
        }
        v[i] = r;
      }
      *reinterpret_cast<uint4*>(orow + c0) = pack16<DT>(v);
    }
  }
}

// image + mixed (constant / map) conditioning channels -> NHWC DT, reflect padded (generalises vk_pack_input)
template <typename DT>
__global__ void pack_input_mixed_kernel(const float* __restrict__ img, int C, int h, int w, int sf, ExtraSrc src,
                                        DT* __restrict__ out, int N, int ld) {
  const int Hp = src.hp, Wp = src.wp;
  const long long total = static_cast<long long>(N) * Hp * Wp;
  const int E = src.ec + src.em;
  constexpr int V = 16 / int(sizeof(DT));
  for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < total;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = int(p % Wp), y = int((p / Wp) % Hp), n = int(p / (static_cast<long long>(Wp) * Hp));
    const int iy = reflect_far(y, src.hh) / sf, ix = reflect_far(x, src.ww) / sf;
    float ex[kMaxE];
    load_extra(src, n, y, x, ex);
    DT* o = out + p * ld;
    for (int c0 = 0; c0 < ld; c0 += V) {
      float v[V];
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int c = c0 + i;
        float r = 0.f;
        if (c < C) {
          r = __ldg(img + ((static_cast<long long>(n) * C + c) * h + iy) * w + ix);
        } else if (c - C < E) {
#pragma unroll
          for (int e = 0; e < kMaxE; ++e)
            if (e == c - C) r = ex[e];
        }
        v[i] = r;
      }
      *reinterpret_cast<uint4*>(o + c0) = pack16<DT>(v);
    }
  }
}

int grid_for_px(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  const long long cap = 148LL * 16;
  return int(std::max(1LL, std::min(b, cap)));
}

template <typename DT, int C2P>
int launch_apply(const vk_sft_apply_args* a, const ExtraSrc& src, const SftW& wt, cudaStream_t st) {
  constexpr int C1P = C2P / 2;
  const int smem = int(sizeof(float)) * (C1P * kMaxE + C1P + C2P * C1P + C2P + 2 * (wt.c * C2P + wt.c));
  auto kern = sft_apply_kernel<DT, C2P>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return int(e);
  }
  const long long total = static_cast<long long>(a->n) * a->h * a->w;
  kern<<<grid_for_px(total, 128), 128, smem, st>>>(reinterpret_cast<const DT*>(a->x), reinterpret_cast<DT*>(a->out),
                                                  a->ld, a->n, a->h, a->w, src, wt, a->alpha, a->round_tf32);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return int(cudaGetLastError());
}

template <typename DT>
int dispatch_apply(const vk_sft_apply_args* a, const ExtraSrc& src, const SftW& wt, cudaStream_t st) {
  const int c2p = (wt.c2 + 7) / 8 * 8;
  switch (c2p) {
    case 8: return launch_apply<DT, 8>(a, src, wt, st);
    case 16: return launch_apply<DT, 16>(a, src, wt, st);
    case 24: return launch_apply<DT, 24>(a, src, wt, st);
    case 32: return launch_apply<DT, 32>(a, src, wt, st);
    case 40: return launch_apply<DT, 40>(a, src, wt, st);
    case 48: return launch_apply<DT, 48>(a, src, wt, st);
    case 56: return launch_apply<DT, 56>(a, src, wt, st);
    case 64: return launch_apply<DT, 64>(a, src, wt, st);
    case 72: return launch_apply<DT, 72>(a, src, wt, st);
    default: return VK_E_UNSUPPORTED;
  }
}


// ---------------------------------------------------------------------------------------------------------
// backward of the per-pixel modulation  a = lrelu(p),  p = x * mul + add,  (mul, add) = AttLayer(e)
// (autograd of networks/AttResUNet.py:27-32,54-58).  The producing dgrad kernel has already applied lrelu'(p), so
// g = dL/dp arrives.  One thread per pixel:
//   gx   = g * mul (+ resid)                                        -> gx   [n][h][w][ld]   (dL/dx)
//   dm   = g * x * mul * (1 - mul)   (through the sigmoid)          -> dm   [n][h][w][ld]
//   df2  = Wm^T dm + Wa^T g ; dq2 = df2 * lrelu'(q2)                -> f2, dq2 [n][h][w][ld2]
//   df1  = W2^T dq2         ; dq1 = df1 * lrelu'(q1)                -> f1, dq1 [n][h][w][ld1]
//   de   = W1^T dq1                                                 -> scattered (atomicAdd) onto the conditioning
//          sources: d_cst [n][ec], d_map [n][em][eh][ew] (sqrt chain rule applied), ev [n][h][w][16] = e
// The parameter gradients are then four pixel-K GEMMs on the tensor cores (vk_conv_wgrad, VK_CONV1X1):
//   gWm = dm^T f2, gbm = sum dm;  gWa = g^T f2, gba = sum g;  gW2 = dq2^T f1, gb2 = sum dq2;  gW1 = dq1^T ev, gb1 = sum dq1.
// ---------------------------------------------------------------------------------------------------------
struct SftBwdOut {
  void *gx, *dm, *f2, *dq2, *f1, *dq1, *ev;
  int ld2, ld1;
  float *d_cst, *d_map;
};

__device__ __forceinline__ void scatter_extra_grad(const ExtraSrc& s, int n, int Y, int X, const float (&de)[kMaxE],
                                                   const float (&ex)[kMaxE], float* d_cst, float* d_map) {
  const int ey = reflect_far(Y, s.hh) / s.esf, exx = reflect_far(X, s.ww) / s.esf;
#pragma unroll
  for (int e = 0; e < kMaxE; ++e) {
    if (e >= s.ec + s.em) break;
    float g = de[e];
    if (s.sqrt_mask & (1u << e)) g *= 0.5f / fmaxf(ex[e], 1e-20f);          // ex = sqrt(raw): d sqrt = 1 / (2 sqrt)
    if (e < s.ec) {
      if (d_cst != nullptr) atomicAdd(d_cst + static_cast<long long>(n) * s.ec + e, g);
    } else if (d_map != nullptr) {
      atomicAdd(d_map + ((static_cast<long long>(n) * s.em + (e - s.ec)) * s.eh + ey) * s.ew + exx, g);
    }
  }
}

template <typename DT>
__device__ __forceinline__ void store_vec_row(DT* dst, int ld, const float* v, int nvalid) {
  constexpr int V = 16 / int(sizeof(DT));
  for (int c0 = 0; c0 < ld; c0 += V) {
    float t[V];
#pragma unroll
    for (int i = 0; i < V; ++i) t[i] = (c0 + i < nvalid) ? v[c0 + i] : 0.f;
    *reinterpret_cast<uint4*>(dst + c0) = pack16<DT>(t);
  }
}

template <typename DT, int C2P>
__global__ void __launch_bounds__(128)
sft_apply_bwd_kernel(const DT* __restrict__ g, const DT* __restrict__ x, const DT* __restrict__ resid, int ld, int N, int h,
                     int w, ExtraSrc src, SftW wt, float alpha, int rnd, SftBwdOut o) {
  extern __shared__ __align__(16) float smem_f[];
  const int E = src.ec + src.em;
  stage_weights<C2P>(smem_f, wt, E);
  __syncthreads();
  constexpr int C1P = C2P / 2;
  constexpr int V = 16 / int(sizeof(DT));
  const float* w1 = smem_f;
  const float* w2 = w1 + C1P * kMaxE + C1P;
  const float* wm = smem_f + C1P * kMaxE + C1P + C2P * C1P + C2P;
  const float* bm = wm + wt.c * C2P;
  const float* wa = bm + wt.c;
  const long long total = static_cast<long long>(N) * h * w;
  const int sy = src.hp / h, sx = src.wp / w;
  for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < total;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int px = int(p % w), py = int((p / w) % h), n = int(p / (static_cast<long long>(w) * h));
    float ex[kMaxE], f1[C1P], f2[C2P], df2[C2P];
    load_extra(src, n, py * sy, px * sx, ex);
    mlp_hidden<C2P>(smem_f, ex, alpha, f1, f2);
#pragma unroll
    for (int j = 0; j < C2P; ++j) df2[j] = 0.f;
    const DT* gr = g + p * ld;
    const DT* xr = x + p * ld;
    DT* gxr = reinterpret_cast<DT*>(o.gx) + p * ld;
    DT* dmr = reinterpret_cast<DT*>(o.dm) + p * ld;
    for (int c0 = 0; c0 < ld; c0 += V) {
      float gv[V], xv[V], rv[V], ogx[V], odm[V];
      unpack16<DT>(*reinterpret_cast<const uint4*>(gr + c0), gv);
      unpack16<DT>(*reinterpret_cast<const uint4*>(xr + c0), xv);
      if (resid != nullptr) unpack16<DT>(*reinterpret_cast<const uint4*>(resid + p * ld + c0), rv);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int c = c0 + i;
        float vgx = 0.f, vdm = 0.f;
        if (c < wt.c) {
          float am = bm[c];
          const float4* wm4 = reinterpret_cast<const float4*>(wm + c * C2P);
#pragma unroll
          for (int j4 = 0; j4 < C2P / 4; ++j4) {
            const float4 m = wm4[j4];
            am = fmaf(m.x, f2[4 * j4], am), am = fmaf(m.y, f2[4 * j4 + 1], am);
            am = fmaf(m.z, f2[4 * j4 + 2], am), am = fmaf(m.w, f2[4 * j4 + 3], am);
          }
          const float mul = 1.f / (1.f + __expf(-am));
          const float gg = gv[i];
          vgx = gg * mul + (resid != nullptr ? rv[i] : 0.f);
          vdm = gg * xv[i] * mul * (1.f - mul);
          // the GEMMs that form the parameter gradients read dm / g in the storage type: back-propagate the same values
          float dmq = vdm;
          if (sizeof(DT) == 2) dmq = __bfloat162float(__float2bfloat16_rn(vdm));
          const float4* wa4 = reinterpret_cast<const float4*>(wa + c * C2P);
#pragma unroll
          for (int j4 = 0; j4 < C2P / 4; ++j4) {
            const float4 m = wm4[j4], a = wa4[j4];
            df2[4 * j4] = fmaf(m.x, dmq, fmaf(a.x, gg, df2[4 * j4]));
            df2[4 * j4 + 1] = fmaf(m.y, dmq, fmaf(a.y, gg, df2[4 * j4 + 1]));
            df2[4 * j4 + 2] = fmaf(m.z, dmq, fmaf(a.z, gg, df2[4 * j4 + 2]));
            df2[4 * j4 + 3] = fmaf(m.w, dmq, fmaf(a.w, gg, df2[4 * j4 + 3]));
          }
        }
        ogx[i] = vgx, odm[i] = vdm;
      }
      *reinterpret_cast<uint4*>(gxr + c0) = pack16<DT>(ogx);
      *reinterpret_cast<uint4*>(dmr + c0) = pack16<DT>(odm);
    }
    // through the two hidden layers
    float dq1[C1P], de[kMaxE];
#pragma unroll
    for (int j = 0; j < C2P; ++j) df2[j] *= (f2[j] > 0.f ? 1.f : alpha);           // dq2
#pragma unroll
    for (int i = 0; i < C1P; ++i) {
      float a = 0.f;
#pragma unroll
      for (int j = 0; j < C2P; ++j) a = fmaf(w2[j * C1P + i], df2[j], a);
      dq1[i] = a * (f1[i] > 0.f ? 1.f : alpha);
    }
#pragma unroll
    for (int e = 0; e < kMaxE; ++e) {
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < C1P; ++i) a = fmaf(w1[i * kMaxE + e], dq1[i], a);
      de[e] = a;
    }
    store_vec_row<DT>(reinterpret_cast<DT*>(o.f2) + p * o.ld2, o.ld2, f2, wt.c2);
    store_vec_row<DT>(reinterpret_cast<DT*>(o.dq2) + p * o.ld2, o.ld2, df2, wt.c2);
    store_vec_row<DT>(reinterpret_cast<DT*>(o.f1) + p * o.ld1, o.ld1, f1, wt.c1);
    store_vec_row<DT>(reinterpret_cast<DT*>(o.dq1) + p * o.ld1, o.ld1, dq1, wt.c1);
    store_vec_row<DT>(reinterpret_cast<DT*>(o.ev) + p * 16, 16, ex, E);
    scatter_extra_grad(src, n, py * sy, px * sx, de, ex, o.d_cst, o.d_map);
  }
}

// gradient of the head convolution's input w.r.t. its conditioning channels (channels [C, C + E) of the packed input,
// networks/AttResUNet.py:152-153), folded back through the reflect padding / nearest up-sampling / sqrt
template <typename DT>
__global__ void extra_head_grad_kernel(const DT* __restrict__ g, int ld, int C, int N, ExtraSrc src, float* d_cst,
                                       float* d_map) {
  const long long total = static_cast<long long>(N) * src.hp * src.wp;
  const int E = src.ec + src.em;
  for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < total;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = int(p % src.wp), y = int((p / src.wp) % src.hp), n = int(p / (static_cast<long long>(src.wp) * src.hp));
    float ex[kMaxE], de[kMaxE];
    load_extra(src, n, y, x, ex);
#pragma unroll
    for (int e = 0; e < kMaxE; ++e) de[e] = e < E ? float(g[p * ld + C + e]) : 0.f;
    scatter_extra_grad(src, n, y, x, de, ex, d_cst, d_map);
  }
}

template <typename DT, int C2P>
int launch_apply_bwd(const vk_sft_apply_bwd_args* a, const ExtraSrc& src, const SftW& wt, cudaStream_t st) {
  constexpr int C1P = C2P / 2;
  const int smem = int(sizeof(float)) * (C1P * kMaxE + C1P + C2P * C1P + C2P + 2 * (wt.c * C2P + wt.c));
  auto kern = sft_apply_bwd_kernel<DT, C2P>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return int(e);
  }
  SftBwdOut o{a->gx, a->dm, a->f2, a->dq2, a->f1, a->dq1, a->ev, a->ld2, a->ld1, a->d_cst, a->d_map};
  const long long total = static_cast<long long>(a->n) * a->h * a->w;
  kern<<<grid_for_px(total, 128), 128, smem, st>>>(reinterpret_cast<const DT*>(a->g), reinterpret_cast<const DT*>(a->x),
                                                  reinterpret_cast<const DT*>(a->resid), a->ld, a->n, a->h, a->w, src, wt,
                                                  a->alpha, 0, o);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return int(cudaGetLastError());
}

template <typename DT>
int dispatch_apply_bwd(const vk_sft_apply_bwd_args* a, const ExtraSrc& src, const SftW& wt, cudaStream_t st) {
  switch ((wt.c2 + 7) / 8 * 8) {
    case 8: return launch_apply_bwd<DT, 8>(a, src, wt, st);
    case 16: return launch_apply_bwd<DT, 16>(a, src, wt, st);
    case 24: return launch_apply_bwd<DT, 24>(a, src, wt, st);
    case 32: return launch_apply_bwd<DT, 32>(a, src, wt, st);
    case 40: return launch_apply_bwd<DT, 40>(a, src, wt, st);
    case 48: return launch_apply_bwd<DT, 48>(a, src, wt, st);
    case 56: return launch_apply_bwd<DT, 56>(a, src, wt, st);
    case 64: return launch_apply_bwd<DT, 64>(a, src, wt, st);
    case 72: return launch_apply_bwd<DT, 72>(a, src, wt, st);
    default: return VK_E_UNSUPPORTED;
  }
}

bool fill_src(const vk_extra_src* e, ExtraSrc* s) {
  if (e->ec < 0 || e->em < 0 || e->ec + e->em > kMaxE) return false;
  if (e->ec > 0 && e->cst == nullptr) return false;
  if (e->em > 0 && (e->map == nullptr || e->eh <= 0 || e->ew <= 0 || e->esf <= 0)) return false;
  if (e->hh <= 0 || e->ww <= 0 || e->hp < e->hh || e->wp < e->ww) return false;
  if (e->hp > 2 * e->hh - 1 || e->wp > 2 * e->ww - 1) return false;
  if (e->em > 0 && ((e->hh + e->esf - 1) / e->esf > e->eh || (e->ww + e->esf - 1) / e->esf > e->ew)) return false;
  s->cst = e->cst, s->map = e->map, s->ec = e->ec, s->em = e->em, s->eh = e->eh, s->ew = e->ew;
  s->esf = e->esf > 0 ? e->esf : 1, s->sqrt_mask = e->sqrt_mask;
  s->hh = e->hh, s->ww = e->ww, s->hp = e->hp, s->wp = e->wp;
  return true;
}
}  // namespace
}  // namespace vk

using namespace vk;

extern "C" int vk_sft_apply(const vk_sft_apply_args* a, void* stream) {
  if (a == nullptr || a->x == nullptr || a->out == nullptr) return VK_E_BADARG;
  if (a->n <= 0 || a->h <= 0 || a->w <= 0 || a->c <= 0 || a->c > a->ld) return VK_E_BADARG;
  const int esize = a->dtype == VK_BF16 ? 2 : 4;
  if ((a->dtype != VK_BF16 && a->dtype != VK_TF32) || (a->ld * esize) % 16) return VK_E_BADARG;
  if (!a->w1 || !a->b1 || !a->w2 || !a->b2 || !a->wm || !a->bm || !a->wa || !a->ba) return VK_E_BADARG;
  if (a->c1 <= 0 || a->c2 <= 0 || 2 * a->c1 > (a->c2 + 7) / 8 * 8) return VK_E_UNSUPPORTED;
  ExtraSrc src;
  if (!fill_src(&a->extra, &src)) return VK_E_BADARG;
  if (src.hp % a->h || src.wp % a->w) return VK_E_BADARG;       // levels of the U-Net divide the padded size
  SftW wt{a->w1, a->b1, a->w2, a->b2, a->wm, a->bm, a->wa, a->ba, a->c1, a->c2, a->c};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->dtype == VK_BF16) return dispatch_apply<__nv_bfloat16>(a, src, wt, st);
  return dispatch_apply<float>(a, src, wt, st);
}

extern "C" int vk_pack_input_mixed(int32_t dtype, const float* img, int32_t n, int32_t c, int32_t h, int32_t w,
                                   int32_t sf, const vk_extra_src* extra, void* out, int32_t ld, void* stream) {
  if (img == nullptr || out == nullptr || extra == nullptr || n <= 0 || c <= 0 || sf <= 0) return VK_E_BADARG;
  ExtraSrc src;
  if (!fill_src(extra, &src)) return VK_E_BADARG;
  if (src.hh != h * sf || src.ww != w * sf || c + src.ec + src.em > ld) return VK_E_BADARG;
  const long long total = static_cast<long long>(n) * src.hp * src.wp;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == VK_BF16)
    pack_input_mixed_kernel<__nv_bfloat16><<<grid_for_px(total, 256), 256, 0, st>>>(
        img, c, h, w, sf, src, reinterpret_cast<__nv_bfloat16*>(out), n, ld);
  else if (dtype == VK_TF32)
    pack_input_mixed_kernel<float><<<grid_for_px(total, 256), 256, 0, st>>>(img, c, h, w, sf, src,
                                                                           reinterpret_cast<float*>(out), n, ld);
  else
    return VK_E_BADARG;
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return int(cudaGetLastError());
}


extern "C" int vk_sft_apply_bwd(const vk_sft_apply_bwd_args* a, void* stream) {
  if (a == nullptr || !a->g || !a->x || !a->gx || !a->dm || !a->f2 || !a->dq2 || !a->f1 || !a->dq1 || !a->ev) return VK_E_BADARG;
  if (a->n <= 0 || a->h <= 0 || a->w <= 0 || a->c <= 0 || a->c > a->ld) return VK_E_BADARG;
  const int esize = a->dtype == VK_BF16 ? 2 : 4;
  if ((a->dtype != VK_BF16 && a->dtype != VK_TF32) || (a->ld * esize) % 16 || (a->ld2 * esize) % 16 || (a->ld1 * esize) % 16)
    return VK_E_BADARG;
  if (!a->w1 || !a->b1 || !a->w2 || !a->b2 || !a->wm || !a->bm || !a->wa || !a->ba) return VK_E_BADARG;
  if (a->c1 <= 0 || a->c2 <= 0 || a->c1 > a->ld1 || a->c2 > a->ld2 || 2 * a->c1 > (a->c2 + 7) / 8 * 8) return VK_E_UNSUPPORTED;
  ExtraSrc src;
  if (!fill_src(&a->extra, &src)) return VK_E_BADARG;
  if (src.hp % a->h || src.wp % a->w) return VK_E_BADARG;
  SftW wt{a->w1, a->b1, a->w2, a->b2, a->wm, a->bm, a->wa, a->ba, a->c1, a->c2, a->c};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->dtype == VK_BF16) return dispatch_apply_bwd<__nv_bfloat16>(a, src, wt, st);
  return dispatch_apply_bwd<float>(a, src, wt, st);
}

extern "C" int vk_extra_head_grad(int32_t dtype, const void* g, int32_t n, int32_t ld, int32_t c, const vk_extra_src* extra,
                                  float* d_cst, float* d_map, void* stream) {
  if (g == nullptr || extra == nullptr || n <= 0 || c < 0) return VK_E_BADARG;
  ExtraSrc src;
  if (!fill_src(extra, &src)) return VK_E_BADARG;
  if (c + src.ec + src.em > ld) return VK_E_BADARG;
  const long long total = static_cast<long long>(n) * src.hp * src.wp;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == VK_BF16)
    extra_head_grad_kernel<__nv_bfloat16><<<grid_for_px(total, 256), 256, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(g), ld, c, n, src, d_cst, d_map);
  else if (dtype == VK_TF32)
    extra_head_grad_kernel<float><<<grid_for_px(total, 256), 256, 0, st>>>(reinterpret_cast<const float*>(g), ld, c, n, src,
                                                                          d_cst, d_map);
  else
    return VK_E_BADARG;
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return int(cudaGetLastError());
}

extern "C" uint32_t vk_sizeof_sft_apply_bwd_args(void) { return uint32_t(sizeof(vk_sft_apply_bwd_args)); }

extern "C" uint32_t vk_sizeof_sft_apply_args(void) { return uint32_t(sizeof(vk_sft_apply_args)); }
