// vk_conv_v2.cuh — persistent implicit-GEMM convolution for sm_100a (second generation of
// vk_conv_igemm.cuh; same maths, same C ABI entry point, selected by the host heuristics).
//
// What changed against v1 (measured on B200, profiles/r01_*):
//   * v1 serialised mainloop and epilogue inside one CTA per tile group and paid TMEM allocation,
//     barrier set-up and the pipeline fill once per group.  v2 is PERSISTENT: one CTA per SM walks
//     the job list, the fp32 accumulators are double-buffered in TMEM (2 x P tiles x N columns), so
//     the epilogue of job j overlaps the TMA/MMA mainloop of job j+1.
//   * 3x3 stride-1 convolutions load ONE halo slab per (tile, K chunk): a (tw+2) x (th+2) pixel box
//     with tw = 8.  A tile row is then exactly one 8-row UMMA core-matrix group, so each of the 9
//     filter taps is the same slab read through a descriptor whose start address is shifted by
//     (r * slab_w + s) rows and whose group pitch (SBO) is slab_w rows — no im2col, and the slab is
//     fetched from L2 once instead of three times (the swizzle is a function of the absolute
//     shared-memory address, so row-shifted descriptors stay consistent; tools/probe `shift`).
//   * A (activation slabs) and B (weights) move through two independent mbarrier rings, so a weight
//     block of `NT` taps can be much smaller than the slab it is multiplied with.
//   * The epilogue no longer issues per-thread 32-byte global accesses (32 cache lines per warp
//     instruction).  Each epilogue warp owns a 32-pixel x `ecols`-channel item: residual / mask
//     arrive by TMA into a private swizzled staging buffer, results leave by TMA store.
//
// Warp roles (13 warps): 0 = A producer, 1-3 = B producers, 4-11 = epilogue (two warps per TMEM lane
// quarter), 12 = MMA issuer and TMEM owner.
#pragma once
#include <type_traits>

#include "vk_conv_igemm.cuh"

namespace vk {

struct ConvV2Load {
  int dx, dy;   // A box origin = tile origin * a_stride + (dx, dy)
  int tap0;     // weight tap of (bi = 0, t = 0); tap = tap0 + bi * NT + t
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
  int nb;                    // B items per A item (1, 3 or 9)
  uint32_t a_off16[9];       // [tap = bi * NT + t]: A descriptor start offset (bytes >> 4) inside the box
  int a_box_bytes;           // smem bytes reserved per A box (multiple of 1024)
  int a_tx_bytes;            // bytes one A box transfers
  int a_sbo;                 // pitch of 8-row groups of the A view (bytes)
  int b_tx_bytes;            // bytes one B item transfers (NT * n_cta * chunk)
  int a_slot_bytes, b_slot_bytes;
  int a_stages, b_stages;
  // ---- jobs ----
  int P;                     // pixel tiles per job
  int n_cta, n_blocks;       // GEMM N per job, N blocks
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
  int ecb;                   // bytes per staged row (64 or 32)
  int ecols;                 // channels per item
  int n_ech;                 // items per tile = n_cta / ecols
  int ebx, eby;              // 32-pixel sub-box of a warp: ebx x eby pixels
  int epi_warp_bytes;        // staging bytes per epilogue warp
  int off_r, off_k, off_o1, off_o2;
  int epi_base;              // byte offset of the staging region in dynamic smem
  int b_base;                // byte offset of the B ring
  // EPI_NCHW_F32 (direct stores)
  void* out1_ptr;
  const void* resid_ptr;
  int act_expclamp;
  float clamp_lo, clamp_hi;
  int crop_h, crop_w;
  long long* timing;         // optional int64[grid][16] stall counters (debug), or null
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

constexpr int kV2Threads = 416;
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

template <typename DT, int kChunkBytes, int kNT>
__global__ void __launch_bounds__(kV2Threads, 1)
conv_v2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ ConvV2Maps emaps, const __grid_constant__ ConvV2Params prm) {
  constexpr bool kTF32 = DTraits<DT>::kTF32;
  constexpr int kElemBytes = sizeof(DT);
  constexpr int kChunkElems = kChunkBytes / kElemBytes;
  constexpr int kMmasPerChunk = kChunkBytes / 32;
  constexpr uint32_t kLayout = layout_type_for_swizzle(kChunkBytes);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_a[kV2MaxStages], empty_a[kV2MaxStages];
  __shared__ __align__(8) uint64_t full_b[kV2MaxStages], empty_b[kV2MaxStages];
  __shared__ __align__(8) uint64_t tmem_full[2], tmem_empty[2];
  __shared__ __align__(8) uint64_t in_bar[8];
  __shared__ uint32_t tmem_base_slot;
  __shared__ float bias_s[1024];

  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + prm.b_base;
  uint8_t* smem_e = smem + prm.epi_base;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int P = prm.P;
  const int tiles_per_img = prm.tiles_x * prm.tiles_y;
  const int n_a_items = prm.n_loads * prm.k_chunks;
  const bool prof = prm.timing != nullptr;
  long long* const tslot = prof ? prm.timing + blockIdx.x * 16 : nullptr;
  const long long t_start = prof ? clock64() : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kV2MaxStages; ++s) {
      mbar_init(&full_a[s], 1), mbar_init(&empty_a[s], 1);
      mbar_init(&full_b[s], 1), mbar_init(&empty_b[s], 1);
      mbar_init(&in_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) mbar_init(&tmem_full[s], 1), mbar_init(&tmem_empty[s], 8);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 12) {
    tmem_alloc(&tmem_base_slot, prm.tmem_cols);
    tmem_relinquish();
  }
  if (warp >= 4 && warp < 12) {
    for (int i = threadIdx.x - 128; i < prm.wrows; i += 256) {
      const int c = i % prm.cq;
      bias_s[i] = (prm.bias != nullptr && c < prm.cout) ? __ldg(prm.bias + c) : 0.f;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ===================== A producer =====================
    if (elect_one()) {
      int sa = 0;
      uint32_t ph = 0;
      long long w_empty = 0;
      for (int job = blockIdx.x; job < prm.n_jobs; job += gridDim.x) {
        const int group = job / prm.n_blocks;
        const int tile0 = group * P;
        const int nvalid = min(P, prm.n_tiles - tile0);
        int timg[4], tx0[4], ty0[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const int t = tile0 + (p < nvalid ? p : 0);
          const int img = t / tiles_per_img;
          const int r = t - img * tiles_per_img;
          const int ty = r / prm.tiles_x, tx = r - ty * prm.tiles_x;
          timg[p] = img, tx0[p] = (tx << prm.tw_log2) * prm.a_stride, ty0[p] = ty * prm.th * prm.a_stride;
        }
        for (int l = 0; l < prm.n_loads; ++l) {
          const int dx = prm.loads[l].dx, dy = prm.loads[l].dy;
          for (int c = 0; c < prm.k_chunks; ++c) {
            mbar_wait_t(&empty_a[sa], ph ^ 1, prof, w_empty);
            mbar_arrive_expect_tx(&full_a[sa], nvalid * prm.a_tx_bytes);
            uint8_t* dst = smem_a + sa * prm.a_slot_bytes;
#pragma unroll
            for (int p = 0; p < 4; ++p)
              if (p < nvalid)
                tma_load_4d(dst + p * prm.a_box_bytes, &tmap_a, &full_a[sa], c * kChunkElems, tx0[p] + dx, ty0[p] + dy,
                            timg[p]);
            if (++sa == prm.a_stages) sa = 0, ph ^= 1;
          }
        }
      }
      if (prof) tslot[5] = w_empty, tslot[11] = clock64() - t_start;
    }
  } else if (warp < 4) {
    // ===================== B producers (B items round-robin over 3 warps) =====================
    const int pw = warp - 1;
    if (elect_one()) {
      int sb = 0, turn = 0;
      uint32_t ph = 0;
      long long w_empty = 0;
      for (int job = blockIdx.x; job < prm.n_jobs; job += gridDim.x) {
        const int n0 = (job % prm.n_blocks) * prm.n_cta;
        for (int l = 0; l < prm.n_loads; ++l) {
          const int tap0 = prm.loads[l].tap0;
          for (int c = 0; c < prm.k_chunks; ++c) {
            for (int bi = 0; bi < prm.nb; ++bi) {
              if (turn == pw) {
                mbar_wait_t(&empty_b[sb], ph ^ 1, prof, w_empty);
                mbar_arrive_expect_tx(&full_b[sb], prm.b_tx_bytes);
                tma_load_3d(smem_b + sb * prm.b_slot_bytes, &tmap_b, &full_b[sb], c * kChunkElems, n0, tap0 + bi * kNT);
              }
              if (++turn == kV2BProducers) turn = 0;
              if (++sb == prm.b_stages) sb = 0, ph ^= 1;
            }
          }
        }
      }
      if (prof && pw == 0) tslot[6] = w_empty, tslot[12] = clock64() - t_start;
    }
  } else if (warp == 12) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc(DTraits<DT>::kFmt, 128, prm.n_cta, 0, 0);
    const uint64_t desc_hi = make_smem_desc(0, 16, prm.a_sbo, kLayout) & 0xFFFFFFFF00000000ull;
    const uint64_t desc_hi_b = make_smem_desc(0, 16, 8 * kChunkBytes, kLayout) & 0xFFFFFFFF00000000ull;
    const uint32_t lbo_lo = 1u << 16;
    const uint32_t a0_16 = ((smem_u32(smem_a) & 0x3FFFFu) >> 4) | lbo_lo;
    const uint32_t b0_16 = ((smem_u32(smem_b) & 0x3FFFFu) >> 4) | lbo_lo;
    const uint32_t aslot16 = uint32_t(prm.a_slot_bytes) >> 4, bslot16 = uint32_t(prm.b_slot_bytes) >> 4;
    const uint32_t box16 = uint32_t(prm.a_box_bytes) >> 4;
    const uint32_t btap16 = uint32_t(prm.n_cta * kChunkBytes) >> 4;
    const uint32_t acc_stride = prm.acc_stride;
    int sa = 0, sb = 0, as = 0;
    uint32_t pha = 0, phb = 0, phacc = 0;
    uint32_t a_lo = a0_16, b_lo = b0_16;
    long long w_fa = 0, w_fb = 0, w_te = 0;
    for (int job = blockIdx.x; job < prm.n_jobs; job += gridDim.x) {
      const int tile0 = (job / prm.n_blocks) * P;
      const int nvalid = min(P, prm.n_tiles - tile0);
      mbar_wait_t(&tmem_empty[as], phacc ^ 1, prof, w_te);
      tc_fence_after_sync();
      const uint32_t d0 = tmem_base + uint32_t(as * P) * acc_stride;
      uint32_t accum = 0;
      for (int ai = 0; ai < n_a_items; ++ai) {
        mbar_wait_t(&full_a[sa], pha, prof, w_fa);
#pragma unroll
        for (int bi = 0; bi < 9; ++bi) {
          if (bi * kNT < 9 && bi < prm.nb) {
            mbar_wait_t(&full_b[sb], phb, prof, w_fb);
            tc_fence_after_sync();
            if (leader) {
              uint32_t ap = a_lo, dp = d0;
              for (int p = 0; p < nvalid; ++p, ap += box16, dp += acc_stride) {
                // tap outer, K step inner: the accumulation order (load, chunk, tap, k) depends on neither the
                // taps-per-stage nor the tiles-per-job picked by the host -> bit-identical across batch sizes
#pragma unroll
                for (int t = 0; t < kNT; ++t) {
#pragma unroll
                  for (int k = 0; k < kMmasPerChunk; ++k) {
                    umma_ss<kTF32>(dp, desc_hi | (ap + prm.a_off16[(bi * kNT + t) % 9] + 2 * k),
                                   desc_hi_b | (b_lo + t * btap16 + 2 * k), idesc, (k == 0 && t == 0) ? accum : 1u);
                  }
                }
              }
              umma_commit(&empty_b[sb]);
            }
            __syncwarp();
            accum = 1;
            b_lo += bslot16;
            if (++sb == prm.b_stages) sb = 0, phb ^= 1, b_lo = b0_16;
          }
        }
        if (leader) umma_commit(&empty_a[sa]);
        __syncwarp();
        a_lo += aslot16;
        if (++sa == prm.a_stages) sa = 0, pha ^= 1, a_lo = a0_16;
      }
      if (leader) umma_commit(&tmem_full[as]);
      __syncwarp();
      if (++as == 2) as = 0, phacc ^= 1;
    }
    if (prof && leader) tslot[1] = w_fa, tslot[2] = w_fb, tslot[3] = w_te, tslot[4] = clock64() - t_start;
  } else {
    // ===================== epilogue (warps 4..11) =====================
    const int ew = warp - 4;
    const int q4 = warp & 3;                   // TMEM lane quarter of this warp
    const int half = ew >> 2;
    const int tw_mask = (1 << prm.tw_log2) - 1;
    const int sub_x = (q4 * 32) & tw_mask, sub_y = (q4 * 32) >> prm.tw_log2;   // origin of the warp's 32 pixels
    const int row = q4 * 32 + lane;
    const int tyy = row >> prm.tw_log2, txx = row & tw_mask;
    const uint32_t stg = smem_u32(smem_e + ew * prm.epi_warp_bytes);
    const uint32_t buf_r = stg + prm.off_r, buf_k = stg + prm.off_k, buf_o1 = stg + prm.off_o1,
                   buf_o2 = stg + prm.off_o2;
    const int ecb = prm.ecb, ecols = prm.ecols, n_ech = prm.n_ech;
    const bool has_in = prm.epi == EPI_STD && (prm.has_resid || prm.has_mask);
    const uint32_t in_bytes = uint32_t(32 * ecb) * uint32_t((prm.has_resid ? 1 : 0) + (prm.has_mask ? 1 : 0));
    const float alpha = prm.alpha;
    uint64_t* my_bar = &in_bar[ew];
    uint32_t in_ph = 0;
    int as = 0;
    uint32_t phacc = 0;
    const bool lead = lane == 0;
    long long w_tf = 0, w_in = 0, w_rd = 0;

    for (int job = blockIdx.x; job < prm.n_jobs; job += gridDim.x) {
      const int group = job / prm.n_blocks;
      const int n0 = (job - group * prm.n_blocks) * prm.n_cta;
      const int qi = n0 / prm.cq;
      const int cq0 = n0 - qi * prm.cq;
      const int tile0 = group * P;
      const int nvalid = min(P, prm.n_tiles - tile0);
      const uint32_t acc0 = tmem_base + (uint32_t(q4 * 32) << 16) + uint32_t(as * P) * prm.acc_stride;

      if (prm.epi == EPI_STD) {
        const int items = nvalid * n_ech;
        // item -> (tile, channel chunk) -> TMA coordinates of this warp's sub-box
        auto coords = [&](int idx, int& c0, int& x, int& y, int& img, int& p, int& kc) {
          p = idx / n_ech;
          kc = idx - p * n_ech;
          const int t = tile0 + p;
          img = t / tiles_per_img;
          const int r = t - img * tiles_per_img;
          const int ty = r / prm.tiles_x, tx = r - ty * prm.tiles_x;
          x = (tx << prm.tw_log2) + sub_x, y = ty * prm.th + sub_y;
          c0 = cq0 + kc * ecols;
        };
        auto issue_in = [&](int idx) {
          int c0, x, y, img, p, kc;
          coords(idx, c0, x, y, img, p, kc);
          mbar_arrive_expect_tx(my_bar, in_bytes);
          if (prm.has_resid)
            tma_load_4d(reinterpret_cast<void*>(smem_e + ew * prm.epi_warp_bytes + prm.off_r), &emaps.resid[qi], my_bar,
                        c0, x, y, img);
          if (prm.has_mask)
            tma_load_4d(reinterpret_cast<void*>(smem_e + ew * prm.epi_warp_bytes + prm.off_k), &emaps.mask, my_bar, c0,
                        x, y, img);
        };
        int idx = half;
        if (has_in && idx < items && lead) issue_in(idx);
        mbar_wait_t(&tmem_full[as], phacc, prof, w_tf);
        tc_fence_after_sync();
        for (; idx < items; idx += 2) {
          int c0, x, y, img, p, kc;
          coords(idx, c0, x, y, img, p, kc);
          const uint32_t taddr = acc0 + uint32_t(p) * prm.acc_stride + uint32_t(kc * ecols);
          // one item = 32 pixels x ecols channels; a staged row is ecb bytes = nchunks 16-byte chunks
          constexpr int CPC = 16 / kElemBytes;            // channels per 16-byte chunk
          const int nchunks = ecb >> 4;                   // 4 or 2
          uint4 rraw[4], kraw[4];
          if (has_in) {
            mbar_wait_t(my_bar, in_ph, prof, w_in);
            in_ph ^= 1;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (j < nchunks) {
                if (prm.has_resid) rraw[j] = lds128(stage_addr(buf_r, lane, ecb, j));
                if (prm.has_mask) kraw[j] = lds128(stage_addr(buf_k, lane, ecb, j));
              }
            }
            __syncwarp();
            if (idx + 2 < items && lead) issue_in(idx + 2);
          }
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
          if (lead) {                             // previous stores have finished reading the out buffers
            if (prof) {
              const long long t0 = clock64();
              bulk_wait_read0();
              w_rd += clock64() - t0;
            } else {
              bulk_wait_read0();
            }
          }
          __syncwarp();
          const float* bias_p = bias_s + n0 + kc * ecols;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j < nchunks) {
              float v[CPC], rv[CPC], mv[CPC];
              if constexpr (kTF32) {
                rv[0] = __uint_as_float(rraw[j].x), rv[1] = __uint_as_float(rraw[j].y);
                rv[2] = __uint_as_float(rraw[j].z), rv[3] = __uint_as_float(rraw[j].w);
                mv[0] = __uint_as_float(kraw[j].x), mv[1] = __uint_as_float(kraw[j].y);
                mv[2] = __uint_as_float(kraw[j].z), mv[3] = __uint_as_float(kraw[j].w);
              } else {
                const uint32_t rw[4] = {rraw[j].x, rraw[j].y, rraw[j].z, rraw[j].w};
                const uint32_t kw[4] = {kraw[j].x, kraw[j].y, kraw[j].z, kraw[j].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  rv[2 * i] = __uint_as_float(rw[i] << 16), rv[2 * i + 1] = __uint_as_float(rw[i] & 0xFFFF0000u);
                  mv[2 * i] = __uint_as_float(kw[i] << 16), mv[2 * i + 1] = __uint_as_float(kw[i] & 0xFFFF0000u);
                }
              }
#pragma unroll
              for (int i = 0; i < CPC; ++i) {
                const int col = j * CPC + i;
                float x = __uint_as_float(acc[col >> 4][col & 15]) + bias_p[col];
                if (prm.has_mask) x *= (mv[i] > 0.f ? 1.f : alpha);
                if (prm.has_resid) x += rv[i];
                v[i] = x;
              }
              auto pack = [&](const float (&f)[CPC]) -> uint4 {
                if constexpr (kTF32) {
                  return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]),
                                    __float_as_uint(f[3]));
                } else {
                  uint32_t w[4];
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
                    w[i] = *reinterpret_cast<uint32_t*>(&h);
                  }
                  return make_uint4(w[0], w[1], w[2], w[3]);
                }
              };
              if (prm.has_out1) sts128(stage_addr(buf_o1, lane, ecb, j), pack(v));
              if (prm.has_out2) {
#pragma unroll
                for (int i = 0; i < CPC; ++i) {
                  v[i] = lrelu(v[i], alpha);
                  if (kTF32 && prm.round_out2) v[i] = round_tf32(v[i]);
                }
                sts128(stage_addr(buf_o2, lane, ecb, j), pack(v));
              }
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lead) {
            if (prm.has_out1)
              tma_store_4d(&emaps.out1[qi], smem_e + ew * prm.epi_warp_bytes + prm.off_o1, c0, x, y, img);
            if (prm.has_out2)
              tma_store_4d(&emaps.out2[qi], smem_e + ew * prm.epi_warp_bytes + prm.off_o2, c0, x, y, img);
            bulk_commit();
          }
        }
      } else {
        // EPI_NCHW_F32: n_cta == 16 columns, direct stores (consecutive lanes = consecutive pixels of a row)
        mbar_wait(&tmem_full[as], phacc);
        tc_fence_after_sync();
        float* const o1 = reinterpret_cast<float*>(prm.out1_ptr);
        const float* const rs = reinterpret_cast<const float*>(prm.resid_ptr);
        const long long plane = static_cast<long long>(prm.crop_h) * prm.crop_w;
        for (int p = half; p < nvalid; p += 2) {
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
      if (lead) mbar_arrive(&tmem_empty[as]);
      if (++as == 2) as = 0, phacc ^= 1;
    }
    if (lead) bulk_wait0();                      // all stores of this warp have completed
    if (prof && lead && ew == 0) tslot[7] = w_tf, tslot[8] = w_in, tslot[9] = w_rd, tslot[10] = clock64() - t_start;
  }

  tc_fence_before_sync();
  __syncthreads();
  if (prof && threadIdx.x == 0) tslot[0] = clock64() - t_start;
  if (warp == 12) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, prm.tmem_cols);
  }
}

}  // namespace vk
