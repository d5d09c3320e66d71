// vk_conv_v2.cuh — persistent implicit-GEMM convolution for sm_100a: serves every dense convolution on
// the VIRNet hot path in both directions (reference call sites: networks/AttResUNet.py:43,46,67,80,117-119,139,
// networks/DnCNN.py:22-29, networks/KNet.py:32-34,49).  vk_conv_igemm.cuh ("v1") is kept as the fallback for
// shapes the host heuristics decline.
//
// GEMM view: M = 128 output pixels (an 8 x 16 patch of one image), N = output channels of the job, K = taps x
// input channels; activations NHWC, weights pre-packed K-major [tap][Cout][Cin].
//
//   * PERSISTENT: one CTA per SM walks the job list; fp32 accumulators are double-buffered in TMEM
//     (2 x P tiles x N columns <= 512), so the epilogue of job j overlaps the TMA/MMA mainloop of job j+1.
//   * CTA PAIRS (kPair, tcgen05 cta_group::2): a cluster of two CTAs walks the same job list; each CTA owns its
//     own pixel tiles and stages HALF of the weight rows; the leader issues M = 256 MMAs.  TMA loads of both CTAs
//     complete on the leader's barriers, commits are multicast, the peer signals "accumulator drained" remotely.
//   * SLAB mode (3x3 stride 1): ONE halo slab per (tile, K chunk), a (8+2) x (16+2) pixel box.  A tile row is
//     exactly one 8-row UMMA core-matrix group, so each of the 9 taps is the same slab read through a descriptor
//     shifted by (r * slab_w + s) rows with group pitch (SBO) slab_w rows — no im2col, one L2 fetch per input
//     pixel (the swizzle is a function of the absolute smem address; tools/probe `shift`).
//   * FULL-K mode (kFullK: stride-2, transposed, 1x1, 2x2 kinds): an A item is one box with a group of K chunks,
//     a weight item one tap with the same chunks; a stride-2 conv reads the four parity phase images of its input
//     as strided boxes whose taps are row-shifted views.
//   * A (activations) and B (weights) move through two independent mbarrier rings; weights of a one-block layer
//     that fit the ring stay resident after the first job.
//   * EPILOGUE by TMA: two groups of four warps own (128-pixel tile x 32-channel chunk) items; residual / mask
//     tiles arrive by TMA into a swizzled staging buffer (prefetched one item ahead), results leave by ONE TMA
//     store per output tensor; the maths is specialised at compile time per tensor combination.
//
// Warp roles (13 warps): 0 = A producer, 1-3 = B producers, 4-11 = epilogue (one warp per TMEM lane quarter in
// each of the two groups), 12 = MMA issuer and TMEM owner.  Optional per-role stall counters (prm.timing) feed
// tools/v2_timing.py.
#pragma once
#include <type_traits>

#include "vk_conv_igemm.cuh"

namespace vk {

struct ConvV2Load {
  int dx, dy;          // A box origin = tile origin * a_stride + (dx, dy)
  int tap0;            // slab mode: weight tap of (bi = 0, t = 0); tap = tap0 + bi * NT + t
  // full-K mode (strided / transposed / 1x1 kinds): this load's box feeds `nb` single-tap weight items
  int nb;
  int tap[4];          // weight tap of item bi
  uint32_t aoff16[4];  // A descriptor start offset of item bi inside the box (bytes >> 4)
};

struct ConvV2Params {
  // ---- M tiling ----
  int n_img, oh, ow;
  int tiles_x, tiles_y, n_tiles;
  int tw_log2, th;
  int a_stride;
  // ---- K loop ----
  int n_loads;
  ConvV2Load loads[9];
  int k_chunks;
  int full_k;                // 1: an A item is one load x one GROUP of kg K chunks (kg boxes per tile) and a B item is
                             // one tap x the same chunks; 0 (3x3 s1 slab): A item = one chunk, B item = NT taps
  int kg;                    // K chunks per group (full_k mode)
  int nb;                    // B items per A item (1, 3 or 9)
  uint32_t a_off16[9];       // [tap = bi * NT + t]: A descriptor start offset (bytes >> 4) inside the box
  int a_box_bytes;           // smem bytes reserved per A box (multiple of 1024)
  int a_tx_bytes;            // bytes one A box transfers
  int a_sbo;                 // pitch of 8-row groups of the A view (bytes)
  int b_tx_bytes;            // bytes one B item transfers (NT * n_cta * chunk)
  int a_slot_bytes, b_slot_bytes;
  int a_stages, b_stages;
  int b_resident;            // 1: the weight ring holds ALL items of a job and n_blocks == 1 -> loaded once per CTA
  // ---- jobs ----
  int P;                     // pixel tiles per job (1, 2 or 4; <= epilogue groups)
  int p_log2;
  int n_cta, n_blocks;       // GEMM N per job, N blocks
  int phase_jobs;            // 1 (merged stride-2 dgrad): the N-block index of a job is a PHASE — it selects the load
                             // (tap views), the epilogue quadrant maps; the weight rows always start at 0
  int n_jobs;
  int acc_stride;            // TMEM columns between accumulators
  int tmem_cols;
  // ---- epilogue ----
  int epi;
  int cq;                    // channels per sub-pixel quadrant (== wrows when no depth-to-space)
  int wrows;
  int cout;
  float alpha;
  int round_out2;
  int has_resid, has_mask, has_out1, has_out2;
  const float* bias;
  const float* sft_mul;      // per-sample modulation of out2 (fp32 [n][sft_ld]) or null
  const float* sft_add;
  int sft_ld;
  int ecb;                   // bytes per staged row (64 or 32)
  int ecols;                 // channels per item
  int n_ech;                 // items per tile = n_cta / ecols
  int ebx, eby;              // 32-pixel sub-box of a warp: ebx x eby pixels
  int epi_warp_bytes;        // staging bytes per epilogue warp
  int off_r, off_k, off_o1, off_o2;
  int epi_base;              // byte offset of the staging region in dynamic smem
  int b_base;                // byte offset of the B ring
  // fast epilogue of the specialised slab kernels: residual / mask rows are read straight from global memory
  const void* mask_ptr;      // NHWC [n][oh][ow][ldo_e] (same layout as the outputs)
  int ldo_e;                 // channel pitch (elements) of the epilogue tensors
  int fast_epi;              // 1: staging area = 2 x (out1, out2) buffers per group, no input staging
  // EPI_NCHW_F32 (direct stores)
  void* out1_ptr;
  const void* resid_ptr;
  int act_expclamp;
  float clamp_lo, clamp_hi;
  int crop_h, crop_w;
  long long* timing;         // optional int64[grid][24] stall counters (debug), or null
};

struct ConvV2Maps {
  CUtensorMap out1[4], out2[4], resid[4], mask;
};

// mbar_wait that adds the stalled cycles to `acc` when profiling is on
__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity, bool prof, long long& acc) {
  if (prof) {
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    acc += clock64() - t0;
  } else {
    mbar_wait(bar, parity);
  }
}

// 4 producer warps + kEpiGroups x 4 epilogue warps + 1 MMA warp.
// (measured: 3 or 4 groups per CTA are slower than 2 — the register cap spills and the larger staging area
// shrinks the weight ring; profiles/r01_v2_*)
constexpr int v2_epi_groups(bool pair) { return pair ? 2 : 2; }
constexpr int v2_threads(bool pair) { return (4 + 4 * v2_epi_groups(pair) + 1) * 32; }
constexpr int kV2MaxStages = 8;
constexpr int kV2BProducers = 3;

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2,
                                             int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_l2_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}

// 16-byte chunk `j` of staged row `row` (rows of `ecb` bytes, TMA swizzle of the same width)
__device__ __forceinline__ uint32_t stage_addr(uint32_t unit, int row, int ecb, int j) {
  const int swz = ecb == 64 ? ((row >> 1) & 3) : ((row >> 2) & 1);
  return unit + uint32_t(row * ecb + ((j ^ swz) << 4));
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// NC consecutive channels of one staged row <-> fp32 registers; c_byte = byte offset of the first channel
template <int NC>
__device__ __forceinline__ void stage_load(uint32_t unit, int row, int ecb, int c_byte, float (&v)[NC],
                                           const __nv_bfloat16*) {
#pragma unroll
  for (int j = 0; j < NC / 8; ++j) {
    const uint4 a = lds128(stage_addr(unit, row, ecb, (c_byte >> 4) + j));
    const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[8 * j + 2 * i] = __uint_as_float(w[i] << 16);
      v[8 * j + 2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
}
template <int NC>
__device__ __forceinline__ void stage_load(uint32_t unit, int row, int ecb, int c_byte, float (&v)[NC], const float*) {
#pragma unroll
  for (int j = 0; j < NC / 4; ++j) {
    const uint4 a = lds128(stage_addr(unit, row, ecb, (c_byte >> 4) + j));
    v[4 * j] = __uint_as_float(a.x), v[4 * j + 1] = __uint_as_float(a.y);
    v[4 * j + 2] = __uint_as_float(a.z), v[4 * j + 3] = __uint_as_float(a.w);
  }
}
template <int NC>
__device__ __forceinline__ void stage_store(uint32_t unit, int row, int ecb, int c_byte, const float (&v)[NC],
                                            const __nv_bfloat16*) {
#pragma unroll
  for (int j = 0; j < NC / 8; ++j) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[8 * j + 2 * i], v[8 * j + 2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    sts128(stage_addr(unit, row, ecb, (c_byte >> 4) + j), make_uint4(w[0], w[1], w[2], w[3]));
  }
}
template <int NC>
__device__ __forceinline__ void stage_store(uint32_t unit, int row, int ecb, int c_byte, const float (&v)[NC],
                                            const float*) {
#pragma unroll
  for (int j = 0; j < NC / 4; ++j)
    sts128(stage_addr(unit, row, ecb, (c_byte >> 4) + j),
           make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
                      __float_as_uint(v[4 * j + 3])));
}


// ---------------------------------------------------------------------------
// Epilogue maths of one item row: ecb bytes = NCH 16-byte chunks of CPC channels each.
//   v = acc (+bias, already added); v *= lrelu'(mask); v += resid; out1 = v; out2 = lrelu(v)
// Compile-time flags keep the per-channel instruction count minimal: the epilogue warps share the issue
// slots of the SM with nothing else that matters, and their instruction count is what bounds small-N layers.
// ---------------------------------------------------------------------------
template <typename DT>
__device__ __forceinline__ void unpack_chunk(const uint4& raw, float (&f)[16 / sizeof(DT)]) {
  if constexpr (sizeof(DT) == 4) {
    f[0] = __uint_as_float(raw.x), f[1] = __uint_as_float(raw.y), f[2] = __uint_as_float(raw.z), f[3] = __uint_as_float(raw.w);
  } else {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) f[2 * i] = __uint_as_float(w[i] << 16), f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}
template <typename DT>
__device__ __forceinline__ uint4 pack_chunk(const float (&f)[16 / sizeof(DT)]) {
  if constexpr (sizeof(DT) == 4) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
  } else {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
}
template <typename DT, bool kMask, bool kResid, bool kOut1, bool kOut2, int kEcb>
__device__ __forceinline__ void epi_item_math(const uint32_t (&acc)[2][16], const uint4 (&rraw)[4], const uint4 (&kraw)[4],
                                              float alpha, bool rnd, uint32_t a_o1, uint32_t a_o2, int swz) {
  constexpr int CPC = 16 / int(sizeof(DT));
  constexpr int NCH = kEcb / 16;
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    float v[CPC];
#pragma unroll
    for (int i = 0; i < CPC; ++i) v[i] = __uint_as_float(acc[(j * CPC + i) >> 4][(j * CPC + i) & 15]);
    if constexpr (kMask) {
      float m[CPC];
      unpack_chunk<DT>(kraw[j], m);
#pragma unroll
      for (int i = 0; i < CPC; ++i) v[i] = m[i] > 0.f ? v[i] : v[i] * alpha;
    }
    if constexpr (kResid) {
      float r[CPC];
      unpack_chunk<DT>(rraw[j], r);
#pragma unroll
      for (int i = 0; i < CPC; ++i) v[i] += r[i];
    }
    const uint32_t sw = uint32_t((j ^ swz) << 4);
    if constexpr (kOut1) sts128(a_o1 + sw, pack_chunk<DT>(v));
    if constexpr (kOut2) {
#pragma unroll
      for (int i = 0; i < CPC; ++i) v[i] = fmaxf(v[i], v[i] * alpha);      // LeakyReLU for 0 < alpha < 1
      if constexpr (sizeof(DT) == 4) {
        if (rnd) {
#pragma unroll
          for (int i = 0; i < CPC; ++i) v[i] = round_tf32(v[i]);
        }
      }
      sts128(a_o2 + sw, pack_chunk<DT>(v));
    }
  }
}

// out2 = lrelu(v * mul + add) with per-(sample, channel) SFT scalars (AttLayer on constant conditioning maps);
// out1 = v.  Runtime flags: only the super-resolution forward path comes through here.
template <typename DT>
__device__ __forceinline__ void epi_item_math_sft(const uint32_t (&acc)[2][16], const uint4 (&rraw)[4], int nch,
                                                  bool has_resid, bool has_out1, const float* __restrict__ mul,
                                                  const float* __restrict__ add, int c_valid, float alpha, bool rnd,
                                                  uint32_t a_o1, uint32_t a_o2, int swz) {
  constexpr int CPC = 16 / int(sizeof(DT));
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (j < nch) {
      float v[CPC], r[CPC];
      unpack_chunk<DT>(rraw[j], r);
#pragma unroll
      for (int i = 0; i < CPC; ++i) {
        v[i] = __uint_as_float(acc[(j * CPC + i) >> 4][(j * CPC + i) & 15]);
        if (has_resid) v[i] += r[i];
      }
      const uint32_t sw = uint32_t((j ^ swz) << 4);
      if (has_out1) sts128(a_o1 + sw, pack_chunk<DT>(v));
#pragma unroll
      for (int i = 0; i < CPC; ++i) {
        const int c = j * CPC + i;
        const float m = c < c_valid ? __ldg(mul + c) : 1.f, a = c < c_valid ? __ldg(add + c) : 0.f;
        v[i] = lrelu(fmaf(v[i], m, a), alpha);
        if (sizeof(DT) == 4 && rnd) v[i] = round_tf32(v[i]);
      }
      sts128(a_o2 + sw, pack_chunk<DT>(v));
    }
  }
}

// kEpi: -1 = every epilogue tensor combination behind a run-time switch; >= 0 = only that combination
// (mask=1, resid=2, out1=4, out2=8, +16 = SFT modulation of out2; 64-byte staging rows, no stall counters) — the
// combinations the large layers of a training step use.  The generic kernel's per-item epilogue walks ~460 instructions scattered
// over 40 KB of SASS (ncu: 41 % of the epilogue warps' stall samples are instruction fetch); the specialised ones keep
// that loop contiguous.
template <typename DT, int kChunkBytes, int kNT, bool kPair, bool kFullK, int kEpi = -1>
__global__ void __launch_bounds__(v2_threads(kPair), 1)
conv_v2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ ConvV2Maps emaps, const __grid_constant__ ConvV2Params prm) {
  constexpr bool kTF32 = DTraits<DT>::kTF32;
  constexpr int kElemBytes = sizeof(DT);
  constexpr int kChunkElems = kChunkBytes / kElemBytes;
  constexpr int kMmasPerChunk = kChunkBytes / 32;
  constexpr uint32_t kLayout = layout_type_for_swizzle(kChunkBytes);
  constexpr int kGroups = v2_epi_groups(kPair);
  constexpr int kMmaWarp = 4 + 4 * kGroups;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_a[kV2MaxStages], empty_a[kV2MaxStages];
  __shared__ __align__(8) uint64_t full_b[kV2MaxStages], empty_b[kV2MaxStages];
  __shared__ __align__(8) uint64_t tmem_full[2], tmem_empty[2];
  __shared__ __align__(8) uint64_t in_bar[8];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float bias_s[1024];
  __shared__ uint32_t lds_sink[128 * kGroups];   // see "inputs consumed" in the staged-input epilogue

  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + prm.b_base;
  uint8_t* smem_e = smem + prm.epi_base;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int P = prm.P;
  const int tiles_per_img = prm.tiles_x * prm.tiles_y;
  // full-K mode: item ai = (load l, chunk group g) -> chunks [c0, c0 + ns); slab mode: item = (load 0, chunk c0), ns = 1
  const int n_kgroups = kFullK ? (prm.k_chunks + prm.kg - 1) / prm.kg : prm.k_chunks;
  const int n_a_items = (prm.phase_jobs ? 1 : prm.n_loads) * n_kgroups;
  const int sub_max = kFullK ? prm.kg : 1;               // A boxes per tile per A item (slot layout)
  auto item_of = [&](int ai, int& l, int& c0, int& ns) {
    l = ai / n_kgroups;
    const int g = ai - l * n_kgroups;
    c0 = kFullK ? g * prm.kg : g;
    ns = kFullK ? min(prm.kg, prm.k_chunks - c0) : 1;
  };
  // pair mode (cta_group::2): the two CTAs of a cluster walk the same job list; CTA `cta_rank` owns tile
  // group 2 * pair_group + cta_rank and half of the weight rows, the leader (rank 0) issues M=256 MMAs
  const uint32_t cta_rank = kPair ? cluster_ctarank() : 0u;
  const bool is_leader = cta_rank == 0;
  const int worker = kPair ? int(blockIdx.x >> 1) : int(blockIdx.x);
  const int n_workers = kPair ? int(gridDim.x >> 1) : int(gridDim.x);
  const int n_groups = (prm.n_tiles + P - 1) / P;
  (void)n_groups;
  const bool prof = kEpi < 0 && prm.timing != nullptr;
  long long* const tslot = prof ? prm.timing + blockIdx.x * 24 : nullptr;
  const long long t_start = prof ? clock64() : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kV2MaxStages; ++s) {
      mbar_init(&full_a[s], 1), mbar_init(&empty_a[s], 1);
      mbar_init(&full_b[s], 1), mbar_init(&empty_b[s], 1);
      mbar_init(&in_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) mbar_init(&tmem_full[s], 1), mbar_init(&tmem_empty[s], (kPair ? 2 : 1) * 4 * kGroups);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == kMmaWarp) {
    if constexpr (kPair) {
      tmem_alloc_2sm(&tmem_base_slot, prm.tmem_cols);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(&tmem_base_slot, prm.tmem_cols);
      tmem_relinquish();
    }
  }
  if (warp >= 4 && warp < kMmaWarp) {
    for (int i = threadIdx.x - 128; i < prm.wrows; i += 128 * kGroups) {
      const int c = i % prm.cq;
      bias_s[i] = (prm.bias != nullptr && c < prm.cout) ? __ldg(prm.bias + c) : 0.f;
    }
  }
  tc_fence_before_sync();
  if constexpr (kPair) {
    cluster_sync_all();          // the peer's barriers must be initialised before anything signals them
  } else {
    __syncthreads();
  }
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;
  // Programmatic dependent launch: the prologue above ran while the previous kernel of the stream was finishing (the only
  // global data it read is the bias vector, a parameter last written by the optimizer kernel, which is never launched
  // programmatically); from here on the kernel reads activations the previous kernel produced and overwrites buffers it
  // may still read, so wait for its completion first.  The next kernel may start its own prologue from now on.
  // Every role that touches global memory waits (A producer, B producers, epilogue warps); the MMA issuer touches none.
  // (Letting the weight producers run ahead of the wait is safe only when a plainly serialised kernel separates
  // vk_pack_weights from the first convolution, as in the engine; it bought 1 % at 2 patches per GPU and nothing
  // elsewhere, so a C-ABI caller is not asked to guarantee that.)
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================== A producer =====================
    if (elect_one()) {
      pdl_wait();
      int sa = 0;
      uint32_t ph = 0;
      long long w_empty = 0;
      for (int job = worker; job < prm.n_jobs; job += n_workers) {
        const int group = kPair ? 2 * (job / prm.n_blocks) + int(cta_rank) : job / prm.n_blocks;
        const int tile0 = group * P;
        // pair mode always moves P boxes per CTA (tiles past the end are fully out of bounds -> zero fill)
        const int nvalid = kPair ? P : min(P, prm.n_tiles - tile0);
        int timg[4], tx0[4], ty0[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const int t = tile0 + (p < nvalid ? p : 0);
          const int img = t / tiles_per_img;
          const int r = t - img * tiles_per_img;
          const int ty = r / prm.tiles_x, tx = r - ty * prm.tiles_x;
          timg[p] = img, tx0[p] = (tx << prm.tw_log2) * prm.a_stride, ty0[p] = ty * prm.th * prm.a_stride;
        }
        for (int ai = 0; ai < n_a_items; ++ai) {
          int l, c0, n_sub;
          item_of(ai, l, c0, n_sub);
          if (prm.phase_jobs) l = job % prm.n_blocks;
          const int dx = prm.loads[l].dx, dy = prm.loads[l].dy;
          mbar_wait_t(&empty_a[sa], ph ^ 1, prof, w_empty);
          uint8_t* dst = smem_a + sa * prm.a_slot_bytes;
          if constexpr (kPair) {
            // both CTAs' boxes complete on the LEADER's barrier, armed by the leader with the pair's bytes
            if (is_leader) mbar_arrive_expect_tx(&full_a[sa], 2 * P * n_sub * prm.a_tx_bytes);
            const uint32_t bar = mapa_shared(smem_u32(&full_a[sa]), 0);
#pragma unroll
            for (int p = 0; p < 4; ++p)
              if (p < nvalid)
                for (int c = 0; c < n_sub; ++c)
                  tma_load_4d_2sm(dst + (p * sub_max + c) * prm.a_box_bytes, &tmap_a, bar, (c0 + c) * kChunkElems,
                                  tx0[p] + dx, ty0[p] + dy, timg[p]);
          } else {
            mbar_arrive_expect_tx(&full_a[sa], nvalid * n_sub * prm.a_tx_bytes);
#pragma unroll
            for (int p = 0; p < 4; ++p)
              if (p < nvalid)
                for (int c = 0; c < n_sub; ++c)
                  tma_load_4d(dst + (p * sub_max + c) * prm.a_box_bytes, &tmap_a, &full_a[sa], (c0 + c) * kChunkElems,
                              tx0[p] + dx, ty0[p] + dy, timg[p]);
          }
          if (++sa == prm.a_stages) sa = 0, ph ^= 1;
        }
      }
      if (prof) tslot[5] = w_empty, tslot[11] = clock64() - t_start;
    }
  } else if (warp < 4) {
    // ===================== B producers (B items round-robin over 3 warps) =====================
    const int pw = warp - 1;
    if (elect_one()) {
      pdl_wait();
      int sb = 0, turn = 0;
      uint32_t ph = 0;
      long long w_empty = 0;
      for (int job = worker; job < prm.n_jobs; job += n_workers) {
        if (prm.b_resident && job != worker) break;    // weights stay in shared memory after the first job
        // pair mode: this CTA supplies rows [n0, n0 + n_cta / 2) of the job's weight block
        const int n0 = (prm.phase_jobs ? 0 : (job % prm.n_blocks) * prm.n_cta) + (kPair ? int(cta_rank) * (prm.n_cta >> 1) : 0);
        for (int ai = 0; ai < n_a_items; ++ai) {
          int l, c, n_sub;
          item_of(ai, l, c, n_sub);
          if (prm.phase_jobs) l = job % prm.n_blocks;
          const int nb_l = kFullK ? prm.loads[l].nb : prm.nb;
          for (int bi = 0; bi < nb_l; ++bi) {
            if (turn == pw) {
              mbar_wait_t(&empty_b[sb], ph ^ 1, prof, w_empty);
              void* dst = smem_b + sb * prm.b_slot_bytes;
              if constexpr (kFullK) {
                // one tap, all K chunks: k_chunks boxes [rows][chunk bytes] one after the other
                const int tap = prm.loads[l].tap[bi];
                const int cb = prm.b_tx_bytes;          // bytes of ONE chunk of a weight item (full-K mode)
                if constexpr (kPair) {
                  if (is_leader) mbar_arrive_expect_tx(&full_b[sb], 2 * n_sub * cb);
                  const uint32_t bar = mapa_shared(smem_u32(&full_b[sb]), 0);
                  for (int cc = 0; cc < n_sub; ++cc)
                    tma_load_3d_2sm(reinterpret_cast<uint8_t*>(dst) + cc * cb, &tmap_b, bar, (c + cc) * kChunkElems, n0,
                                    tap);
                } else {
                  mbar_arrive_expect_tx(&full_b[sb], n_sub * cb);
                  for (int cc = 0; cc < n_sub; ++cc)
                    tma_load_3d(reinterpret_cast<uint8_t*>(dst) + cc * cb, &tmap_b, &full_b[sb], (c + cc) * kChunkElems,
                                n0, tap);
                }
              } else {
                const int tap = prm.loads[l].tap0 + bi * kNT;
                if constexpr (kPair) {
                  if (is_leader) mbar_arrive_expect_tx(&full_b[sb], 2 * prm.b_tx_bytes);
                  tma_load_3d_2sm(dst, &tmap_b, mapa_shared(smem_u32(&full_b[sb]), 0), c * kChunkElems, n0, tap);
                } else {
                  mbar_arrive_expect_tx(&full_b[sb], prm.b_tx_bytes);
                  tma_load_3d(dst, &tmap_b, &full_b[sb], c * kChunkElems, n0, tap);
                }
              }
            }
            if (++turn == kV2BProducers) turn = 0;
            if (++sb == prm.b_stages) sb = 0, ph ^= 1;
          }
        }
      }
      if (prof && pw == 0) tslot[6] = w_empty, tslot[12] = clock64() - t_start;
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (pair mode: leader CTA only) =====================
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc(DTraits<DT>::kFmt, kPair ? 256 : 128, prm.n_cta, 0, 0);
    const uint64_t desc_hi = make_smem_desc(0, 16, prm.a_sbo, kLayout) & 0xFFFFFFFF00000000ull;
    const uint64_t desc_hi_b = make_smem_desc(0, 16, 8 * kChunkBytes, kLayout) & 0xFFFFFFFF00000000ull;
    const uint32_t lbo_lo = 1u << 16;
    const uint32_t a0_16 = ((smem_u32(smem_a) & 0x3FFFFu) >> 4) | lbo_lo;
    const uint32_t b0_16 = ((smem_u32(smem_b) & 0x3FFFFu) >> 4) | lbo_lo;
    const uint32_t aslot16 = uint32_t(prm.a_slot_bytes) >> 4, bslot16 = uint32_t(prm.b_slot_bytes) >> 4;
    const uint32_t box16 = uint32_t(prm.a_box_bytes) >> 4;
    const uint32_t btap16 = uint32_t((kPair ? prm.n_cta >> 1 : prm.n_cta) * kChunkBytes) >> 4;
    const uint32_t bchunk16 = btap16;           // full-K mode: K chunk c of a weight item follows chunk c - 1
    const uint32_t acc_stride = prm.acc_stride;
    int sa = 0, sb = 0, as = 0;
    uint32_t pha = 0, phb = 0, phacc = 0;
    uint32_t a_lo = a0_16, b_lo = b0_16;
    long long w_fa = 0, w_fb = 0, w_te = 0;
    for (int job = worker; job < prm.n_jobs && is_leader; job += n_workers) {
      const int tile0 = (job / prm.n_blocks) * P;
      const int nvalid = kPair ? P : min(P, prm.n_tiles - tile0);
      mbar_wait_t(&tmem_empty[as], phacc ^ 1, prof, w_te);
      tc_fence_after_sync();
      const uint32_t d0 = tmem_base + uint32_t(as * P) * acc_stride;
      uint32_t accum = 0;
      for (int ai = 0; ai < n_a_items; ++ai) {
        mbar_wait_t(&full_a[sa], pha, prof, w_fa);
        int l_, c0_, n_sub;
        item_of(ai, l_, c0_, n_sub);
        if (prm.phase_jobs) l_ = job % prm.n_blocks;
        const int nb_l = kFullK ? prm.loads[l_].nb : prm.nb;
#pragma unroll
        for (int bi = 0; bi < 9; ++bi) {
          if (bi * kNT < 9 && bi < nb_l) {
            if (!prm.b_resident || job == worker) mbar_wait_t(&full_b[sb], phb, prof, w_fb);
            tc_fence_after_sync();
            if (leader) {
              // full-K mode: item bi of load ai reads its tap view at aoff16[bi]; sub-box c is K chunk c
              const uint32_t a_item = a_lo + (kFullK ? prm.loads[l_].aoff16[bi & 3] : 0u);
              uint32_t dp = d0;
              for (int p = 0; p < nvalid; ++p, dp += acc_stride) {
               for (int c = 0; c < n_sub; ++c) {
                const uint32_t ap = a_item + uint32_t(p * sub_max + c) * box16;
                const uint32_t bq = b_lo + uint32_t(c) * bchunk16;
                // tap outer, K step inner: the accumulation order (load, chunk, tap, k) depends on neither the
                // taps-per-stage nor the tiles-per-job picked by the host -> bit-identical across batch sizes
#pragma unroll
                for (int t = 0; t < kNT; ++t) {
#pragma unroll
                  for (int k = 0; k < kMmasPerChunk; ++k) {
                    const uint32_t acc_flag = (k == 0 && t == 0 && c == 0) ? accum : 1u;
                    if constexpr (kPair) {
                      umma_ss_2sm<kTF32>(dp, desc_hi | (ap + prm.a_off16[(bi * kNT + t) % 9] + 2 * k),
                                         desc_hi_b | (bq + t * btap16 + 2 * k), idesc, acc_flag);
                    } else {
                      umma_ss<kTF32>(dp, desc_hi | (ap + prm.a_off16[(bi * kNT + t) % 9] + 2 * k),
                                     desc_hi_b | (bq + t * btap16 + 2 * k), idesc, acc_flag);
                    }
                  }
                }
               }
              }
              if (!prm.b_resident) {
                if constexpr (kPair) umma_commit_2sm(&empty_b[sb]); else umma_commit(&empty_b[sb]);
              }
            }
            __syncwarp();
            accum = 1;
            b_lo += bslot16;
            if (++sb == prm.b_stages) sb = 0, phb ^= 1, b_lo = b0_16;
          }
        }
        if (leader) {
          if constexpr (kPair) umma_commit_2sm(&empty_a[sa]); else umma_commit(&empty_a[sa]);
        }
        __syncwarp();
        a_lo += aslot16;
        if (++sa == prm.a_stages) sa = 0, pha ^= 1, a_lo = a0_16;
      }
      if (leader) {
        if constexpr (kPair) umma_commit_2sm(&tmem_full[as]); else umma_commit(&tmem_full[as]);
      }
      __syncwarp();
      if (++as == 2) as = 0, phacc ^= 1;
    }
    if (prof && leader) tslot[1] = w_fa, tslot[2] = w_fb, tslot[3] = w_te, tslot[4] = clock64() - t_start;
  } else {
    // ===================== epilogue (warps 4 .. 4 + 4 * kGroups - 1) =====================
    pdl_wait();
    const int ew = warp - 4;
    const int q4 = warp & 3;                   // TMEM lane quarter of this warp
    const int half = ew >> 2;
    const int tw_mask = (1 << prm.tw_log2) - 1;
    // the four warps of a `half` form a group that owns whole (tile, channel chunk) items: their 4 x 32 TMEM
    // lanes are the 128 pixels of the tile, staged as rows of one buffer and moved by ONE TMA op per tensor
    const int row = q4 * 32 + lane;
    const int tyy = row >> prm.tw_log2, txx = row & tw_mask;
    uint8_t* const stg_p = smem_e + half * prm.epi_warp_bytes;
    const uint32_t stg = smem_u32(stg_p);
    const int grp_bar = 1 + half;                         // named barrier of the group (128 threads)
    auto group_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(grp_bar) : "memory"); };
    const bool glead = (q4 == 0) && (lane == 0);          // the group's TMA-load-issuing thread
    const bool slead = (q4 == 1) && (lane == 0);          // ... and its TMA-store-issuing thread (bulk groups are per thread)
    // specialised kernels (kEpi >= 0) know the tensor combination and the staging row width at compile time
    const bool e_mask = kEpi >= 0 ? (kEpi & 1) != 0 : prm.has_mask != 0;
    const bool e_resid = kEpi >= 0 ? (kEpi & 2) != 0 : prm.has_resid != 0;
    const bool e_out1 = kEpi >= 0 ? (kEpi & 4) != 0 : prm.has_out1 != 0;
    const bool e_out2 = kEpi >= 0 ? (kEpi & 8) != 0 : prm.has_out2 != 0;
    const int ecb = kEpi >= 0 ? 64 : prm.ecb, ecols = kEpi >= 0 ? 64 / kElemBytes : prm.ecols, n_ech = prm.n_ech;
    const int swz = ecb == 64 ? ((row >> 1) & 3) : ((row >> 2) & 1);      // TMA swizzle of this thread's staged row
    const uint32_t row_base = stg + uint32_t(row * ecb);
    const uint32_t off_r = prm.off_r, off_k = prm.off_k, off_o1 = prm.off_o1, off_o2 = prm.off_o2;
    const bool has_in = prm.epi == EPI_STD && (e_resid || e_mask);
    const uint32_t in_bytes = uint32_t(128 * ecb) * uint32_t((e_resid ? 1 : 0) + (e_mask ? 1 : 0));
    const float alpha = prm.alpha;
    uint64_t* my_bar = &in_bar[half];
    uint32_t in_ph = 0;
    int as = 0;
    uint32_t phacc = 0;
    const bool lead = lane == 0;
    long long w_tf = 0, w_in = 0, w_rd = 0, w_ld = 0, w_math = 0, w_sync = 0, w_st = 0, w_misc = 0, w_fence = 0;
    long long tq = 0;
    uint32_t fast_cnt = 0;                     // items this group has staged so far (selects the output buffer)
    (void)fast_cnt;
    auto lap = [&](long long& acc) {
      if (prof) {
        const long long now = clock64();
        acc += now - tq;
        tq = now;
      }
    };

    for (int job = worker; job < prm.n_jobs; job += n_workers) {
      const int jgroup = job / prm.n_blocks;
      const int group = kPair ? 2 * jgroup + int(cta_rank) : jgroup;
      const int n0 = prm.phase_jobs ? 0 : (job - jgroup * prm.n_blocks) * prm.n_cta;
      const int qi = prm.phase_jobs ? job - jgroup * prm.n_blocks : n0 / prm.cq;
      const int cq0 = n0 - (prm.phase_jobs ? 0 : qi * prm.cq);
      const int tile0 = group * P;
      const int nvalid = max(0, min(P, prm.n_tiles - tile0));
      const uint32_t acc0 = tmem_base + (uint32_t(q4 * 32) << 16) + uint32_t(as * P) * prm.acc_stride;

      if constexpr (kEpi >= 0 && !kFullK && (kEpi & 3) == 0) {
        // ---------------- fast epilogue of the specialised slab kernels WITHOUT input tensors ----------------
        // (measured, profiles/r02_epilogue_experiments.txt: reading residual / mask rows straight from global memory into
        //  registers, or storing results straight from registers, both LOSE to the TMA-staged path — 64-byte rows with a
        //  192-byte pitch make every warp access 32 partial sectors; only the double-buffered output staging below pays)
        // * residual / mask rows come straight from global memory into registers (64 B per thread and tensor), after an
        //   L2 prefetch (TMA prefetch, no smem, no barrier) issued two items ahead — no input staging, no wait on a
        //   staging barrier, nothing to protect with a group barrier;
        // * the output staging is double-buffered: the only barrier per item is the one that publishes the staged
        //   tile to the thread that issues the TMA stores, and the store it has to wait for is one item old.
        const int n_items = P * n_ech;
        const int g_items = nvalid > 0 && half < n_items ? (n_items - half + kGroups - 1) / kGroups : 0;
        int tc_img[2], tc_x[2], tc_y[2];
#pragma unroll
        for (int pp = 0; pp < 2; ++pp) {
          const int t = tile0 + pp;
          tc_img[pp] = t / tiles_per_img;
          const int r = t - tc_img[pp] * tiles_per_img;
          const int ty = r / prm.tiles_x, tx = r - ty * prm.tiles_x;
          tc_x[pp] = tx << prm.tw_log2, tc_y[pp] = ty * prm.th;
        }
        constexpr bool kIn = (kEpi & 3) != 0;
        auto prefetch_item = [&](int g) {            // one thread: pull the item's residual / mask tiles into L2
          const int idx = half + g * kGroups;
          const int pp = idx & (P - 1);
          const int c0 = cq0 + (idx >> prm.p_log2) * ecols;
          const int x = pp ? tc_x[1] : tc_x[0], y = pp ? tc_y[1] : tc_y[0], img = pp ? tc_img[1] : tc_img[0];
          if (e_resid) tma_prefetch_l2_4d(&emaps.resid[0], c0, x, y, img);
          if (e_mask) tma_prefetch_l2_4d(&emaps.mask, c0, x, y, img);
        };
        if (kIn && glead) {
          if (g_items > 0) prefetch_item(0);
          if (g_items > 1) prefetch_item(1);
        }
        mbar_wait(&tmem_full[as], phacc);
        tc_fence_after_sync();
        for (int g = 0; g < g_items; ++g) {
          const int idx = half + g * kGroups;
          const int p = idx & (P - 1);
          const int kc = idx >> prm.p_log2;
          const int t_x = p ? tc_x[1] : tc_x[0], t_y = p ? tc_y[1] : tc_y[0], t_img = p ? tc_img[1] : tc_img[0];
          const uint32_t taddr = acc0 + uint32_t(p) * prm.acc_stride + uint32_t(kc * ecols);
          const int c0 = cq0 + kc * ecols;
          // ---- inputs: this thread's pixel row, 4 x 16 B per tensor, issued before the accumulator load ----
          uint4 rraw[4], kraw[4];
          if constexpr (kIn) {
            if (g + 2 < g_items && glead) prefetch_item(g + 2);
            const int py = t_y + tyy, px = t_x + txx;
            const bool ok = t_img < prm.n_img && py < prm.oh && px < prm.ow;
            const long long off = ((static_cast<long long>(t_img) * prm.oh + py) * prm.ow + px) * prm.ldo_e + c0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (e_resid)
                rraw[j] = ok ? ldg_nc_v4(reinterpret_cast<const DT*>(prm.resid_ptr) + off + j * (16 / kElemBytes))
                             : make_uint4(0u, 0u, 0u, 0u);
              if (e_mask)
                kraw[j] = ok ? ldg_nc_v4(reinterpret_cast<const DT*>(prm.mask_ptr) + off + j * (16 / kElemBytes))
                             : make_uint4(0u, 0u, 0u, 0u);
            }
          }
          uint32_t acc[2][16];
          __syncwarp();
          tmem_ld16(taddr, acc[0]);
          if constexpr (kElemBytes == 2) tmem_ld16(taddr + 16, acc[1]);
          tmem_ld_wait();
          if (prm.bias != nullptr) {              // bias: 16-byte broadcast loads, added in place
            const float4* bp = reinterpret_cast<const float4*>(bias_s + n0 + kc * ecols);
#pragma unroll
            for (int i4 = 0; i4 < 16 / kElemBytes; ++i4) {
              const float4 b = bp[i4];
              uint32_t* a4 = &acc[(i4 * 4) >> 4][(i4 * 4) & 15];
              a4[0] = __float_as_uint(__uint_as_float(a4[0]) + b.x), a4[1] = __float_as_uint(__uint_as_float(a4[1]) + b.y);
              a4[2] = __float_as_uint(__uint_as_float(a4[2]) + b.z), a4[3] = __float_as_uint(__uint_as_float(a4[3]) + b.w);
            }
          }
          const uint32_t obuf = uint32_t(fast_cnt & 1) * uint32_t(prm.epi_warp_bytes >> 1);
          const uint32_t a_o1 = row_base + obuf + off_o1, a_o2 = row_base + obuf + off_o2;
          const bool rnd = kTF32 && prm.round_out2;
          if constexpr (kEpi >= 16) {
            epi_item_math_sft<DT>(acc, rraw, 4, (kEpi & 2) != 0, (kEpi & 4) != 0,
                                  prm.sft_mul + static_cast<long long>(t_img) * prm.sft_ld + c0,
                                  prm.sft_add + static_cast<long long>(t_img) * prm.sft_ld + c0, prm.cout - c0, alpha,
                                  rnd, a_o1, a_o2, swz);
          } else {
            epi_item_math<DT, (kEpi & 1) != 0, (kEpi & 2) != 0, (kEpi & 4) != 0, (kEpi & 8) != 0, 64>(acc, rraw, kraw, alpha,
                                                                                                  rnd, a_o1, a_o2, swz);
          }
          fence_proxy_async_smem();
          if (slead) bulk_wait_read0();           // the store that last read the OTHER buffer (one item old) is done
          group_sync();                           // the whole 128-pixel item is staged
          if (slead) {
            if (e_out1) tma_store_4d(&emaps.out1[0], stg_p + obuf + prm.off_o1, c0, t_x, t_y, t_img);
            if (e_out2) tma_store_4d(&emaps.out2[0], stg_p + obuf + prm.off_o2, c0, t_x, t_y, t_img);
            bulk_commit();
          }
          ++fast_cnt;
        }
      } else
      if (prm.epi == EPI_STD) {
        // Items of a job are (channel chunk kc, tile p), p fastest (P is 1 or 2), dealt round-robin to the
        // kGroups groups: group `half` takes items half, half + kGroups, ...  Tiles past the end of the image
        // list decode to an image index >= n_img, i.e. fully out of bounds for TMA: their loads return zeros
        // and their stores write nothing, so no item needs special casing.
        const int n_items = P * n_ech;
        const int g_items = nvalid > 0 && half < n_items ? (n_items - half + kGroups - 1) / kGroups : 0;
        int tc_img[2], tc_x[2], tc_y[2];
#pragma unroll
        for (int pp = 0; pp < 2; ++pp) {
          const int t = tile0 + pp;
          tc_img[pp] = t / tiles_per_img;
          const int r = t - tc_img[pp] * tiles_per_img;
          const int ty = r / prm.tiles_x, tx = r - ty * prm.tiles_x;
          tc_x[pp] = tx << prm.tw_log2, tc_y[pp] = ty * prm.th;
        }
        auto issue_in = [&](int g) {
          const int idx = half + g * kGroups;
          const int pp = idx & (P - 1);
          const int c0 = cq0 + (idx >> prm.p_log2) * ecols;
          const int x = pp ? tc_x[1] : tc_x[0], y = pp ? tc_y[1] : tc_y[0], img = pp ? tc_img[1] : tc_img[0];
          mbar_arrive_expect_tx(my_bar, in_bytes);
          if (e_resid)
            tma_load_4d(reinterpret_cast<void*>(stg_p + prm.off_r), &emaps.resid[qi], my_bar, c0, x, y, img);
          if (e_mask)
            tma_load_4d(reinterpret_cast<void*>(stg_p + prm.off_k), &emaps.mask, my_bar, c0, x, y, img);
        };
        auto prefetch_in = [&](int g) {
          const int idx = half + g * kGroups;
          const int pp = idx & (P - 1);
          const int c0 = cq0 + (idx >> prm.p_log2) * ecols;
          const int x = pp ? tc_x[1] : tc_x[0], y = pp ? tc_y[1] : tc_y[0], img = pp ? tc_img[1] : tc_img[0];
          if (e_resid) tma_prefetch_l2_4d(&emaps.resid[qi], c0, x, y, img);
          if (e_mask) tma_prefetch_l2_4d(&emaps.mask, c0, x, y, img);
        };
        if (has_in && g_items > 0 && glead) {
          issue_in(0);
          if (g_items > 1) prefetch_in(1);
        }
        mbar_wait_t(&tmem_full[as], phacc, prof, w_tf);
        tc_fence_after_sync();
        for (int g = 0; g < g_items; ++g) {
          if (prof) tq = clock64();
          const int idx = half + g * kGroups;
          const int p = idx & (P - 1);
          const int kc = idx >> prm.p_log2;
          const int t_x = p ? tc_x[1] : tc_x[0], t_y = p ? tc_y[1] : tc_y[0], t_img = p ? tc_img[1] : tc_img[0];
          const uint32_t taddr = acc0 + uint32_t(p) * prm.acc_stride + uint32_t(kc * ecols);
          // ---- inputs: this thread's row of the staged residual / mask tiles ----
          uint4 rraw[4], kraw[4];
          if (has_in) {
            mbar_wait_t(my_bar, in_ph, prof, w_in);
            in_ph ^= 1;
#pragma unroll
            uint32_t sink = 0;
            for (int j = 0; j < 4; ++j) {
              if (j < (ecb >> 4)) {
                if (e_resid) rraw[j] = lds128(row_base + off_r + ((j ^ swz) << 4)), sink ^= rraw[j].x ^ rraw[j].w;
                if (e_mask) kraw[j] = lds128(row_base + off_k + ((j ^ swz) << 4)), sink ^= kraw[j].x ^ kraw[j].w;
              }
            }
            // The TMA load of the NEXT item overwrites this staging area right after the group barrier below.  The barrier
            // orders generic-proxy accesses among the four warps, but a shared-memory load that is still queued in the
            // memory pipe when its warp arrives can be overtaken by that async-proxy write (seen as a rare stale
            // 16-byte mask chunk in one quarter-warp of the latest warp; tools/det_check.py).  A store that consumes
            // every loaded register cannot issue before the loads have returned and cannot move across the barrier.
            asm volatile("st.volatile.shared.b32 [%0], %1;" ::"r"(smem_u32(&lds_sink[half * 128 + row])), "r"(sink) : "memory");
          }
          if (slead) bulk_wait_read0();           // previous stores have finished reading the out buffers
          lap(w_misc);
          group_sync();                           // inputs consumed by all 4 warps, out buffers free
          lap(w_sync);
          if (has_in && g + 1 < g_items && glead) issue_in(g + 1);
          if (has_in && g + 2 < g_items && glead) prefetch_in(g + 2);   // pull the item after next into L2
          // ---- accumulators ----
          uint32_t acc[2][16];
          __syncwarp();
          if (ecols >= 16) {
            tmem_ld16(taddr, acc[0]);
            if (ecols == 32) tmem_ld16(taddr + 16, acc[1]);
          } else {
            uint32_t t8[8];
            tmem_ld8(taddr, t8);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[0][i] = t8[i];
          }
          tmem_ld_wait();
          lap(w_ld);
          if (prm.bias != nullptr) {              // bias: 16-byte broadcast loads, added in place
            const float4* bp = reinterpret_cast<const float4*>(bias_s + n0 + kc * ecols);
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
              if (i4 * 4 < ecols) {
                const float4 b = bp[i4];
                uint32_t* a4 = &acc[(i4 * 4) >> 4][(i4 * 4) & 15];
                a4[0] = __float_as_uint(__uint_as_float(a4[0]) + b.x), a4[1] = __float_as_uint(__uint_as_float(a4[1]) + b.y);
                a4[2] = __float_as_uint(__uint_as_float(a4[2]) + b.z), a4[3] = __float_as_uint(__uint_as_float(a4[3]) + b.w);
              }
            }
          }
          // ---- fused epilogue maths, specialised at compile time on the tensors present ----
          const uint32_t a_o1 = row_base + off_o1, a_o2 = row_base + off_o2;
          if constexpr (kEpi >= 16) {             // SFT epilogue (super-resolution forward): out2 = lrelu(v * mul + add)
            const bool rnd = kTF32 && prm.round_out2;
            const int c0 = cq0 + kc * ecols;
            epi_item_math_sft<DT>(acc, rraw, 4, (kEpi & 2) != 0, (kEpi & 4) != 0,
                                  prm.sft_mul + static_cast<long long>(t_img) * prm.sft_ld + c0,
                                  prm.sft_add + static_cast<long long>(t_img) * prm.sft_ld + c0, prm.cout - c0, alpha,
                                  rnd, a_o1, a_o2, swz);
          } else if constexpr (kEpi >= 0) {
            const bool rnd = kTF32 && prm.round_out2;
            epi_item_math<DT, (kEpi & 1) != 0, (kEpi & 2) != 0, (kEpi & 4) != 0, (kEpi & 8) != 0, 64>(acc, rraw, kraw, alpha,
                                                                                                  rnd, a_o1, a_o2, swz);
          } else {
          const int mode = (prm.has_mask ? 1 : 0) | (prm.has_resid ? 2 : 0) | (prm.has_out1 ? 4 : 0) | (prm.has_out2 ? 8 : 0);
          const bool rnd = kTF32 && prm.round_out2;
#define VK_EPI_CASE(M)                                                                                             \
  case M:                                                                                                          \
    if (ecb == 64)                                                                                                 \
      epi_item_math<DT, (M & 1) != 0, (M & 2) != 0, (M & 4) != 0, (M & 8) != 0, 64>(acc, rraw, kraw, alpha, rnd, a_o1, \
                                                                                    a_o2, swz);                    \
    else                                                                                                           \
      epi_item_math<DT, (M & 1) != 0, (M & 2) != 0, (M & 4) != 0, (M & 8) != 0, 32>(acc, rraw, kraw, alpha, rnd, a_o1, \
                                                                                    a_o2, swz);                    \
    break;
          if (prm.sft_mul != nullptr) {
            const int c0 = cq0 + kc * ecols;
            epi_item_math_sft<DT>(acc, rraw, ecb >> 4, prm.has_resid != 0, prm.has_out1 != 0,
                                  prm.sft_mul + static_cast<long long>(t_img) * prm.sft_ld + c0,
                                  prm.sft_add + static_cast<long long>(t_img) * prm.sft_ld + c0, prm.cout - c0, alpha,
                                  rnd, a_o1, a_o2, swz);
          } else
          switch (mode) {
            VK_EPI_CASE(4) VK_EPI_CASE(5) VK_EPI_CASE(6) VK_EPI_CASE(7) VK_EPI_CASE(8) VK_EPI_CASE(12) VK_EPI_CASE(14)
            default: break;                       // the host only launches the combinations above
          }
#undef VK_EPI_CASE
          }
          lap(w_math);
          fence_proxy_async_smem();
          lap(w_fence);
          group_sync();                           // the whole 128-pixel item is staged
          lap(w_sync);
          if (slead) {
            const int c0 = cq0 + kc * ecols;
            if (e_out1) tma_store_4d(&emaps.out1[qi], stg_p + prm.off_o1, c0, t_x, t_y, t_img);
            if (e_out2) tma_store_4d(&emaps.out2[qi], stg_p + prm.off_o2, c0, t_x, t_y, t_img);
            bulk_commit();
          }
          lap(w_st);
        }
      } else {
        // EPI_NCHW_F32: n_cta == 16 columns, direct stores (consecutive lanes = consecutive pixels of a row)
        mbar_wait(&tmem_full[as], phacc);
        tc_fence_after_sync();
        float* const o1 = reinterpret_cast<float*>(prm.out1_ptr);
        const float* const rs = reinterpret_cast<const float*>(prm.resid_ptr);
        const long long plane = static_cast<long long>(prm.crop_h) * prm.crop_w;
        for (int p = half; p < nvalid; p += kGroups) {
          const int t = tile0 + p;
          const int img = t / tiles_per_img;
          const int r = t - img * tiles_per_img;
          const int ty = r / prm.tiles_x, tx = r - ty * prm.tiles_x;
          const int oy = ty * prm.th + tyy, ox = (tx << prm.tw_log2) + txx;
          const bool in_crop = oy < prm.oh && ox < prm.ow && oy < prm.crop_h && ox < prm.crop_w;
          for (int jc = 0; jc < prm.n_cta; jc += 16) {
            uint32_t rr[16];
            __syncwarp();
            tmem_ld16(acc0 + uint32_t(p) * prm.acc_stride + jc, rr);
            tmem_ld_wait();
            const int co = cq0 + jc;
            const int nch = (in_crop && co < prm.cout) ? min(16, prm.cout - co) : 0;
            const long long base = (static_cast<long long>(img) * prm.cout + co) * plane +
                                   static_cast<long long>(oy) * prm.crop_w + ox;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (i < nch) {
                float xv = __uint_as_float(rr[i]) + bias_s[n0 + jc + i];
                if (prm.act_expclamp) xv = expf(fminf(fmaxf(xv, prm.clamp_lo), prm.clamp_hi));
                if (rs != nullptr) xv += __ldg(rs + base + i * plane);
                o1[base + i * plane] = xv;
              }
            }
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lead) {
        if constexpr (kPair) mbar_arrive_cluster_relaxed(mapa_shared(smem_u32(&tmem_empty[as]), 0)); else mbar_arrive(&tmem_empty[as]);
      }
      if (++as == 2) as = 0, phacc ^= 1;
    }
    if (slead) bulk_wait0();                     // all stores of this group have completed
    if (prof && glead && half == 0) tslot[7] = w_tf, tslot[8] = w_in, tslot[9] = w_rd, tslot[10] = clock64() - t_start, tslot[13] = w_ld, tslot[14] = w_math, tslot[15] = w_sync, tslot[16] = w_st, tslot[17] = w_misc, tslot[18] = w_fence;
  }

  tc_fence_before_sync();
  if constexpr (kPair) {
    cluster_sync_all();          // neither CTA may retire while the other can still touch its smem / barriers / TMEM
  } else {
    __syncthreads();
  }
  if (prof && threadIdx.x == 0) tslot[0] = clock64() - t_start;
  if (warp == kMmaWarp) {
    tc_fence_after_sync();
    if constexpr (kPair) tmem_dealloc_2sm(tmem_base, prm.tmem_cols); else tmem_dealloc(tmem_base, prm.tmem_cols);
  }
}

}  // namespace vk
