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
// CTA = 13 warps.  Warp 0 and warps 6-8: TMA producers (one TMA instruction occupies its issuing
// warp for ~800 cycles on B200 — measured, tools/probe — so the boxes of a stage are spread over
// four warps).  Warp 1: MMA issuer (+TMEM owner); its loop is warp-uniform with only the tcgen05
// instructions predicated on one elected lane, so descriptors live in uniform registers and the
// shallow tensor-core queue is never starved by address arithmetic.  Warps 2-5 and 9-12: epilogue
// (TMEM -> registers -> fused epilogue -> global); two warps share each TMEM lane quarter and
// split the 16-column groups between them.
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
  long long* cta_timing;      // optional per-CTA cycle counters (debug), or null
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

constexpr int kConvThreads = 416;
constexpr int kConvProducers = 4;

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
  if (prm.cta_timing != nullptr && threadIdx.x == 0) {
    long long* t = prm.cta_timing + (blockIdx.y * gridDim.x + blockIdx.x) * 8;
    t[0] = clock64();
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    t[6] = smid;
  }

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
  if (warp >= 2 && warp < 6 && prm.bias != nullptr) {
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

  if (warp == 0 || (warp >= 6 && warp <= 8)) {
    // ===================== TMA producers =====================
    // producer `pw` issues the boxes whose running index is == pw (mod kConvProducers); producer 0
    // also arms the barrier (a complete_tx that lands before the expect_tx is legal: the phase
    // cannot complete before producer 0's arrival).
    const int pw = warp == 0 ? 0 : warp - 5;
    if (elect_one()) {
      int it = 0;
      int op = 0;
      long long prod_wait = 0;
      for (int l = 0; l < prm.n_loads; ++l) {
        const ConvLoad& ld = prm.loads[l];
        for (int c = 0; c < prm.k_chunks; ++c, ++it) {
          const int s = it % prm.stages;
          const uint32_t ph = (it / prm.stages) & 1;
          const long long tw0 = clock64();
          mbar_wait(&empty_bar[s], ph ^ 1);
          prod_wait += clock64() - tw0;
          uint8_t* a_s = smem + s * stage_bytes;
          uint8_t* b_s = a_s + P * box_bytes;
          if (pw == 0) mbar_arrive_expect_tx(&full_bar[s], nvalid * box_bytes + ld.ntaps * b_tap_bytes);
          for (int p = 0; p < nvalid; ++p, ++op) {
            if (op % kConvProducers != pw) continue;
            const int t = tile0 + p;
            const int img = t / tiles_per_img;
            const int r = t - img * tiles_per_img;
            const int ty = r / prm.tiles_x, tx = r - ty * prm.tiles_x;
            const int ox0 = tx << prm.tw_log2, oy0 = ty * prm.th;
            tma_load_4d(a_s + p * box_bytes, &tmap_a, &full_bar[s], c * kChunkElems, ox0 * prm.a_stride + ld.dx,
                        oy0 * prm.a_stride + ld.dy, img);
          }
          for (int j = 0; j < ld.ntaps; ++j, ++op) {
            if (op % kConvProducers != pw) continue;
            tma_load_3d(b_s + j * b_tap_bytes, &tmap_b, &full_bar[s], c * kChunkElems, n0, ld.tap[j]);
          }
        }
      }
      if (prm.cta_timing != nullptr && pw == 0)
        prm.cta_timing[(blockIdx.y * gridDim.x + blockIdx.x) * 8 + 3] = prod_wait;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The issuing thread runs on the (slow, in-order) uniform datapath: every instruction on the
    // dependent chain in front of a tcgen05.mma delays it, so descriptors are kept as 32-bit "low
    // words" (start address >> 4 | LBO) advanced by additions only; the high word is constant.
    {
      const bool leader = elect_one();
      const uint32_t idesc = make_idesc(DTraits<DT>::kFmt, 128, prm.n_cta, 0, 0);
      const uint64_t desc_hi = make_smem_desc(0, 16, kSBO, kLayout) & 0xFFFFFFFF00000000ull;
      const uint32_t lbo_lo = 1u << 16;                       // LBO = 16 bytes (unused for swizzled K-major)
      const uint32_t box16 = uint32_t(box_bytes) >> 4, btap16 = uint32_t(b_tap_bytes) >> 4;
      const uint32_t stage16 = uint32_t(stage_bytes) >> 4;
      const uint32_t smem16 = (smem_u32(smem) & 0x3FFFFu) >> 4;
      const uint32_t b_off16 = uint32_t(P) * box16;
      const uint32_t acc_stride = prm.acc_stride;
      int s = 0;
      uint32_t ph = 0;
      uint32_t accum = 0;
      long long mma_wait = 0;
      for (int l = 0; l < prm.n_loads; ++l) {
        const ConvLoad& ld = prm.loads[l];
        const bool three = ld.ntaps > 1;
        const uint32_t ro0 = uint32_t(ld.rowoff[0] * kChunkBytes) >> 4;
        const uint32_t ro1 = uint32_t(ld.rowoff[1] * kChunkBytes) >> 4;
        const uint32_t ro2 = uint32_t(ld.rowoff[2] * kChunkBytes) >> 4;
        for (int c = 0; c < prm.k_chunks; ++c) {
          if (prm.cta_timing != nullptr) {
            const long long tw0 = clock64();
            mbar_wait(&full_bar[s], ph);
            mma_wait += clock64() - tw0;
          } else {
            mbar_wait(&full_bar[s], ph);
          }
          tc_fence_after_sync();
          const uint32_t a_lo = (smem16 + uint32_t(s) * stage16) | lbo_lo;
          const uint32_t b_lo = a_lo + b_off16;
          uint32_t a_p = a_lo, d_p = tmem_base;
          for (int p = 0; p < nvalid; ++p, a_p += box16, d_p += acc_stride) {
            // K step outer, tap inner: the accumulation order (load, channel, tap) does not depend on
            // the chunk width picked by the host heuristics -> bit-identical results across batch sizes
            if (leader) {
#pragma unroll
              for (int k = 0; k < kMmasPerChunk; ++k) {
                umma_ss<kTF32>(d_p, desc_hi | (a_p + ro0 + 2 * k), desc_hi | (b_lo + 2 * k), idesc,
                               (k == 0) ? accum : 1u);
                if (three) {
                  umma_ss<kTF32>(d_p, desc_hi | (a_p + ro1 + 2 * k), desc_hi | (b_lo + btap16 + 2 * k), idesc, 1u);
                  umma_ss<kTF32>(d_p, desc_hi | (a_p + ro2 + 2 * k), desc_hi | (b_lo + 2 * btap16 + 2 * k), idesc, 1u);
                }
              }
            }
          }
          accum = 1;
          if (leader) umma_commit(&empty_bar[s]);   // frees the smem stage once these MMAs retire
          __syncwarp();
          if (++s == prm.stages) s = 0, ph ^= 1;
        }
      }
      if (leader) {
        umma_commit(&tmem_full_bar);      // accumulators complete
        if (prm.cta_timing != nullptr) {
          long long* t = prm.cta_timing + (blockIdx.y * gridDim.x + blockIdx.x) * 8;
          t[4] = mma_wait;
          t[5] = clock64();               // all MMAs issued
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5 and 9..12) =====================
    const int q4 = warp & 3;                    // TMEM lane quarter this warp may read
    const int half = warp >= 9 ? 1 : 0;         // which of the two warps of this quarter
    const int row = q4 * 32 + lane;             // GEMM row == pixel within the tile
    const int tyy = row >> prm.tw_log2, txx = row & tw_mask;
    // everything that needs an integer division is hoisted out of the tile / column loops
    const int us = prm.us;
    const int quad = n0 / prm.cq + prm.quad_base;            // sub-pixel quadrant: constant per CTA
    const int co0 = n0 - (n0 / prm.cq) * prm.cq;             // first channel (within the quadrant) of this CTA
    const int qy = quad / us, qx = quad - qy * us;
    const int ngroups = prm.n_cta >> 4;
    const float alpha = prm.alpha;
    const bool has_bias = prm.bias != nullptr;
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after_sync();
    if (prm.cta_timing != nullptr && threadIdx.x == 64)
      prm.cta_timing[(blockIdx.y * gridDim.x + blockIdx.x) * 8 + 1] = clock64();
    int img = tile0 / tiles_per_img;
    int rt = tile0 - img * tiles_per_img;
    for (int p = 0; p < nvalid; ++p) {
      const int ty = rt / prm.tiles_x, tx = rt - ty * prm.tiles_x;
      const int oy = ty * prm.th + tyy, ox = (tx << prm.tw_log2) + txx;
      const bool pix_ok = (oy < prm.oh) && (ox < prm.ow);
      const uint32_t taddr = tmem_base + (uint32_t(q4 * 32) << 16) + p * prm.acc_stride;
      if (prm.epi == EPI_STD) {
        const int yy = oy * us + qy, xx = ox * us + qx;
        const bool out_ok = pix_ok && yy < prm.out_h && xx < prm.out_w;
        const long long obase = ((static_cast<long long>(img) * prm.out_h + yy) * prm.out_w + xx) * prm.ldo + co0;
        DT* const o1 = reinterpret_cast<DT*>(prm.out1);
        DT* const o2 = reinterpret_cast<DT*>(prm.out2);
        const DT* const rs = reinterpret_cast<const DT*>(prm.resid);
        const DT* const mk = reinterpret_cast<const DT*>(prm.mask);
        // software pipeline: the mask / residual vectors of the next group are in flight while this one is processed
        float m_nx[16], r_nx[16];
        int g = half;
        if (g < ngroups && out_ok && co0 + g * 16 + 16 <= prm.cout) {
          if (mk != nullptr) load16(mk + obase + g * 16, m_nx);
          if (rs != nullptr) load16(rs + obase + g * 16, r_nx);
        }
        for (; g < ngroups; g += 2) {
          const int jc = g * 16;
          uint32_t rr[16];
          __syncwarp();                           // tcgen05.ld is warp-collective (.sync.aligned)
          tmem_ld16(taddr + jc, rr);
          float m[16], rv[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) m[i] = m_nx[i], rv[i] = r_nx[i];
          const int gn = g + 2;
          if (gn < ngroups && out_ok && co0 + gn * 16 + 16 <= prm.cout) {
            if (mk != nullptr) load16(mk + obase + gn * 16, m_nx);
            if (rs != nullptr) load16(rs + obase + gn * 16, r_nx);
          }
          tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rr[i]);
          if (has_bias) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += bias_s[jc + i];
          }
          const int co = co0 + jc;
          const int nch = (out_ok && co < prm.cout) ? min(16, prm.cout - co) : 0;
          const long long off = obase + jc;
          if (nch == 16) {
            if (mk != nullptr) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] *= (m[i] > 0.f ? 1.f : alpha);
            }
            if (rs != nullptr) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += rv[i];
            }
            if (o1 != nullptr) store16(o1 + off, v);
            if (o2 != nullptr) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                v[i] = lrelu(v[i], alpha);
                if (kTF32 && prm.round_out2) v[i] = round_tf32(v[i]);
              }
              store16(o2 + off, v);
            }
          } else if (nch > 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (i < nch) {
                float x = v[i];
                if (mk != nullptr) x *= (to_float(mk[off + i]) > 0.f ? 1.f : alpha);
                if (rs != nullptr) x += to_float(rs[off + i]);
                if (o1 != nullptr) from_float(o1 + off + i, x);
                if (o2 != nullptr) {
                  float y = lrelu(x, alpha);
                  if (kTF32 && prm.round_out2) y = round_tf32(y);
                  from_float(o2 + off + i, y);
                }
              }
            }
          }
        }
      } else {  // EPI_NCHW_F32 (us == 1)
        float* const o1 = reinterpret_cast<float*>(prm.out1);
        const float* const rs = reinterpret_cast<const float*>(prm.resid);
        const long long plane = static_cast<long long>(prm.crop_h) * prm.crop_w;
        const bool in_crop = pix_ok && oy < prm.crop_h && ox < prm.crop_w;
        for (int g = half; g < ngroups; g += 2) {
          const int jc = g * 16;
          uint32_t rr[16];
          __syncwarp();
          tmem_ld16(taddr + jc, rr);
          tmem_ld_wait();
          const int co = co0 + jc;
          const int nch = (in_crop && co < prm.cout) ? min(16, prm.cout - co) : 0;
          const long long base = (static_cast<long long>(img) * prm.cout + co) * plane +
                                 static_cast<long long>(oy) * prm.crop_w + ox;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (i < nch) {
              float x = __uint_as_float(rr[i]) + (has_bias ? bias_s[jc + i] : 0.f);
              if (prm.act_expclamp) x = expf(fminf(fmaxf(x, prm.clamp_lo), prm.clamp_hi));
              if (rs != nullptr) x += __ldg(rs + base + i * plane);
              o1[base + i * plane] = x;
            }
          }
        }
      }
      if (++rt == tiles_per_img) rt = 0, ++img;
    }
    tc_fence_before_sync();
    if (prm.cta_timing != nullptr && threadIdx.x == 64)
      prm.cta_timing[(blockIdx.y * gridDim.x + blockIdx.x) * 8 + 2] = clock64();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, prm.tmem_cols);
  }
}

}  // namespace vk
