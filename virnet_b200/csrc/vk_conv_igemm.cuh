// vk_conv_igemm.cuh — im2col-free implicit-GEMM convolution on tcgen05 (sm_100a).
//
// One kernel serves every dense convolution on the VIRNet hot path
// (reference call sites: networks/AttResUNet.py:43,46,67,80,117-119,139,
// networks/DnCNN.py:22-29, networks/KNet.py:32-34,45,49) in both directions:
//   fprop 3x3 s1   : 3 "slab" loads (one per horizontal tap s), each feeding
//                    the 3 vertical taps from row-shifted views of one box
//   fprop 3x3 s2   : 9 strided loads (TMA elementStrides = 2)
//   dgrad 3x3 s1   : same as fprop with rotated / transposed weights
//   ConvT 2x2 s2   : 1 load (1x1 GEMM to 4*Cout) + depth-to-space epilogue
//
// GEMM view: M = output pixels (128 per tile, TW x TH patch of one image),
// N = output channels, K = taps x input channels.  Activations are NHWC; the
// A operand of tap (r,s) is the TMA box of the input shifted by (r-1,s-1), with
// out-of-bounds zero fill standing in for the padding, so no im2col buffer ever
// exists.  Weights are pre-packed K-major as [tap][Cout][Cin].
//
// CTA = 6 warps: warp 0 TMA producer, warp 1 MMA issuer (+TMEM owner),
// warps 2-5 epilogue (TMEM -> registers -> fused epilogue -> global).
// A CTA owns P pixel tiles x n_cta channels, i.e. P fp32 accumulators of
// 128 lanes x n_cta columns in TMEM, so each weight tile read from L2 is
// reused P times.
#pragma once
#include "vk_common.cuh"

namespace vk {

struct ConvLoad {
  int dx, dy;         // offset added to (tile origin * a_stride) for the box origin
  int ntaps;          // taps fed by this box (1..3)
  int tap[3];         // weight tap index (3rd coordinate of the weight tensor map)
  int rowoff[3];      // first box row (in pixels) of the tap's 128-row A view
};

enum ConvEpilogue : int {
  EPI_STD = 0,        // NHWC DT: v=acc+bias; v*=mask'; v+=resid; out1=v; out2=lrelu(v)
  EPI_NCHW_F32 = 1,   // NCHW fp32: v=acc+bias; [v=exp(clamp(v))]; [v+=resid_nchw]; crop
};

struct ConvIgemmParams {
  // M tiling
  int n_img, oh, ow;          // pixel grid walked by the M tiles
  int tiles_x, tiles_y, n_tiles;
  int tw_log2, th;            // tile = (1<<tw_log2) x th pixels == 128
  int box_rows;               // pixels per A box (bh * tw)
  int a_stride;               // tile origin -> A coordinate multiplier (conv stride)
  int k_chunks;               // channel chunks per load
  int n_loads;
  ConvLoad loads[9];
  int tiles_per_cta;          // P
  int n_cta;                  // GEMM N per CTA (multiple of 16, <= 256)
  int b_taps;                 // max taps per load (sizes the B region of a stage)
  int acc_stride;             // TMEM columns between accumulators
  int tmem_cols;              // power of two >= P * acc_stride
  int stages;
  // epilogue
  int epi;
  int us;                     // depth-to-space factor (1, or 2 for ConvT / stride-2 dgrad phases)
  int quad_base;              // sub-pixel index added to the column-derived quadrant
  int out_h, out_w;           // spatial dims of the EPI_STD output tensors
  int cq;                     // channels per sub-pixel quadrant (== cout when us == 1)
  int cout;                   // valid output channels (per quadrant)
  int ldo;                    // channel pitch of out1/out2/resid/mask (elements)
  float alpha;                // LeakyReLU slope for out2 and for the mask
  int round_out2;             // tf32 mode: round out2 to tf32 (RN) when storing
  const float* bias;          // [cout] fp32 (the parameter itself) or null
  const void* resid;          // DT NHWC (EPI_STD) / fp32 NCHW (EPI_NCHW_F32) or null
  const void* mask;           // DT NHWC: multiply by (mask>0 ? 1 : alpha), or null
  void* out1;                 // or null
  void* out2;                 // or null
  // EPI_NCHW_F32 only
  int act_expclamp;           // 1: v = exp(clamp(v, lo, hi))
  float clamp_lo, clamp_hi;
  int crop_h, crop_w;         // stored region (<= oh, ow); out/resid are [n][cout][crop_h][crop_w]
};

template <typename DT>
struct DTraits;
template <>
struct DTraits<__nv_bfloat16> {
  static constexpr bool kTF32 = false;
  static constexpr uint32_t kFmt = 1;  // BF16
};
template <>
struct DTraits<float> {
  static constexpr bool kTF32 = true;
  static constexpr uint32_t kFmt = 2;  // TF32
};

// 16 consecutive channels of one pixel <-> registers
__device__ __forceinline__ void load16(const __nv_bfloat16* p, float (&v)[16]) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ void load16(const float* p, float (&v)[16]) {
  const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float4 t = __ldg(q + i);
    v[4 * i] = t.x, v[4 * i + 1] = t.y, v[4 * i + 2] = t.z, v[4 * i + 3] = t.w;
  }
}
__device__ __forceinline__ void store16(__nv_bfloat16* p, const float (&v)[16]) {
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(w[0], w[1], w[2], w[3]);
  q[1] = make_uint4(w[4], w[5], w[6], w[7]);
}
__device__ __forceinline__ void store16(float* p, const float (&v)[16]) {
  float4* q = reinterpret_cast<float4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
__device__ __forceinline__ float to_float(__nv_bfloat16 x) { return __bfloat162float(x); }
__device__ __forceinline__ float to_float(float x) { return x; }
__device__ __forceinline__ void from_float(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void from_float(float* p, float v) { *p = v; }

constexpr int kConvThreads = 192;

template <typename DT, int kChunkBytes>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ ConvIgemmParams prm) {
  constexpr bool kTF32 = DTraits<DT>::kTF32;
  constexpr int kElemBytes = sizeof(DT);
  constexpr int kChunkElems = kChunkBytes / kElemBytes;
  constexpr int kMmasPerChunk = kChunkBytes / 32;            // every UMMA consumes 32 B of K
  constexpr uint32_t kLayout = layout_type_for_swizzle(kChunkBytes);
  constexpr uint32_t kSBO = 8 * kChunkBytes;                 // 8-row core-matrix group pitch
  constexpr int kMaxStages = 8;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;
  __shared__ float bias_s[256];

  // 1024-align the dynamic region by hand (the attribute is not honoured for extern arrays)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int P = prm.tiles_per_cta;
  const int tile0 = blockIdx.x * P;
  const int nvalid = min(P, prm.n_tiles - tile0);
  const int n0 = blockIdx.y * prm.n_cta;
  const int box_bytes = prm.box_rows * kChunkBytes;
  const int b_tap_bytes = prm.n_cta * kChunkBytes;
  const int stage_bytes = P * box_bytes + prm.b_taps * b_tap_bytes;
  const int tiles_per_img = prm.tiles_x * prm.tiles_y;
  const int tw_mask = (1 << prm.tw_log2) - 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < prm.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_slot, prm.tmem_cols);
    tmem_relinquish();
  }
  if (warp >= 2 && prm.bias != nullptr) {
    // bias is the raw parameter [cout]; ConvT repeats it for each of the 4 sub-pixel quadrants
    for (int i = threadIdx.x - 64; i < prm.n_cta; i += 128) {
      const int c = (n0 + i) % prm.cq;
      bias_s[i] = c < prm.cout ? __ldg(prm.bias + c) : 0.f;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int it = 0;
      for (int l = 0; l < prm.n_loads; ++l) {
        const ConvLoad& ld = prm.loads[l];
        for (int c = 0; c < prm.k_chunks; ++c, ++it) {
          const int s = it % prm.stages;
          const uint32_t ph = (it / prm.stages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* a_s = smem + s * stage_bytes;
          uint8_t* b_s = a_s + P * box_bytes;
          mbar_arrive_expect_tx(&full_bar[s], nvalid * box_bytes + ld.ntaps * b_tap_bytes);
          for (int p = 0; p < nvalid; ++p) {
            const int t = tile0 + p;
            const int img = t / tiles_per_img;
            const int r = t - img * tiles_per_img;
            const int ty = r / prm.tiles_x, tx = r - ty * prm.tiles_x;
            const int ox0 = tx << prm.tw_log2, oy0 = ty * prm.th;
            tma_load_4d(a_s + p * box_bytes, &tmap_a, &full_bar[s], c * kChunkElems, ox0 * prm.a_stride + ld.dx,
                        oy0 * prm.a_stride + ld.dy, img);
          }
          for (int j = 0; j < ld.ntaps; ++j)
            tma_load_3d(b_s + j * b_tap_bytes, &tmap_b, &full_bar[s], c * kChunkElems, n0, ld.tap[j]);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint32_t idesc = make_idesc(DTraits<DT>::kFmt, 128, prm.n_cta, 0, 0);
      int it = 0;
      uint32_t accum = 0;
      for (int l = 0; l < prm.n_loads; ++l) {
        const ConvLoad& ld = prm.loads[l];
        for (int c = 0; c < prm.k_chunks; ++c, ++it) {
          const int s = it % prm.stages;
          const uint32_t ph = (it / prm.stages) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after_sync();
          const uint32_t a_s = smem_u32(smem + s * stage_bytes);
          const uint32_t b_s = a_s + P * box_bytes;
          for (int p = 0; p < nvalid; ++p) {
            uint32_t acc_p = accum;
            for (int j = 0; j < ld.ntaps; ++j) {
              const uint32_t a_tap = a_s + p * box_bytes + ld.rowoff[j] * kChunkBytes;
              const uint32_t b_tap = b_s + j * b_tap_bytes;
#pragma unroll
              for (int k = 0; k < kMmasPerChunk; ++k) {
                const uint64_t ad = make_smem_desc(a_tap + k * 32, 16, kSBO, kLayout);
                const uint64_t bd = make_smem_desc(b_tap + k * 32, 16, kSBO, kLayout);
                umma_ss<kTF32>(tmem_base + p * prm.acc_stride, ad, bd, idesc, acc_p);
                acc_p = 1;
              }
            }
          }
          accum = 1;
          umma_commit(&empty_bar[s]);   // frees the smem stage once these MMAs retire
        }
      }
      umma_commit(&tmem_full_bar);      // accumulators complete
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q4 = warp & 3;                    // TMEM lane quarter this warp may read
    const int row = q4 * 32 + lane;             // GEMM row == pixel within the tile
    const int tyy = row >> prm.tw_log2, txx = row & tw_mask;
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after_sync();
    const int us = prm.us;
    for (int p = 0; p < nvalid; ++p) {
      const int t = tile0 + p;
      const int img = t / tiles_per_img;
      const int r = t - img * tiles_per_img;
      const int ty = r / prm.tiles_x, tx = r - ty * prm.tiles_x;
      const int oy = ty * prm.th + tyy, ox = (tx << prm.tw_log2) + txx;
      const bool pix_ok = (oy < prm.oh) && (ox < prm.ow);
      const uint32_t taddr = tmem_base + (uint32_t(q4 * 32) << 16) + p * prm.acc_stride;
      for (int jc = 0; jc < prm.n_cta; jc += 16) {
        uint32_t rr[16];
        __syncwarp();                           // tcgen05.ld is warp-collective (.sync.aligned)
        tmem_ld16(taddr + jc, rr);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rr[i]);
        if (prm.bias != nullptr) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += bias_s[jc + i];
        }
        const int cg = n0 + jc;                 // global GEMM column of v[0]
        const int quad_c = cg / prm.cq;
        const int co = cg - quad_c * prm.cq;    // channel within the quadrant
        const int quad = quad_c + prm.quad_base;
        const bool out_ok = pix_ok && (prm.epi != EPI_STD ||
                                       ((oy * us + quad / us) < prm.out_h && (ox * us + quad % us) < prm.out_w));
        const int nch = (out_ok && co < prm.cout) ? min(16, prm.cout - co) : 0;
        if (nch == 0) {
          // nothing to store for this thread (ragged tile edge / channel padding)
        } else if (prm.epi == EPI_STD) {
          const int qy = quad / us, qx = quad - qy * us;
          const int yy = oy * us + qy, xx = ox * us + qx;
          const long long opix = (static_cast<long long>(img) * prm.out_h + yy) * prm.out_w + xx;
          const long long off = opix * prm.ldo + co;
          DT* o1 = reinterpret_cast<DT*>(prm.out1);
          DT* o2 = reinterpret_cast<DT*>(prm.out2);
          const DT* rs = reinterpret_cast<const DT*>(prm.resid);
          const DT* mk = reinterpret_cast<const DT*>(prm.mask);
          if (nch == 16) {
            if (mk != nullptr) {
              float m[16];
              load16(mk + off, m);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] *= (m[i] > 0.f ? 1.f : prm.alpha);
            }
            if (rs != nullptr) {
              float m[16];
              load16(rs + off, m);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += m[i];
            }
            if (o1 != nullptr) store16(o1 + off, v);
            if (o2 != nullptr) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                v[i] = lrelu(v[i], prm.alpha);
                if (kTF32 && prm.round_out2) v[i] = round_tf32(v[i]);
              }
              store16(o2 + off, v);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (i < nch) {
                float x = v[i];
                if (mk != nullptr) x *= (to_float(mk[off + i]) > 0.f ? 1.f : prm.alpha);
                if (rs != nullptr) x += to_float(rs[off + i]);
                if (o1 != nullptr) from_float(o1 + off + i, x);
                if (o2 != nullptr) {
                  float y = lrelu(x, prm.alpha);
                  if (kTF32 && prm.round_out2) y = round_tf32(y);
                  from_float(o2 + off + i, y);
                }
              }
            }
          }
        } else {  // EPI_NCHW_F32
          if (oy < prm.crop_h && ox < prm.crop_w) {
            float* o1 = reinterpret_cast<float*>(prm.out1);
            const float* rs = reinterpret_cast<const float*>(prm.resid);
            const long long plane = static_cast<long long>(prm.crop_h) * prm.crop_w;
            const long long base = (static_cast<long long>(img) * prm.cout + co) * plane +
                                   static_cast<long long>(oy) * prm.crop_w + ox;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (i < nch) {
                float x = v[i];
                if (prm.act_expclamp) x = expf(fminf(fmaxf(x, prm.clamp_lo), prm.clamp_hi));
                if (rs != nullptr) x += __ldg(rs + base + i * plane);
                o1[base + i * plane] = x;
              }
            }
          }
        }
      }
    }
    tc_fence_before_sync();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, prm.tmem_cols);
  }
}

}  // namespace vk
