// vk_wgrad.cuh — weight-gradient of the convolutions as a tcgen05 GEMM whose K
// dimension is the pixel index (sm_100a).
//
//   conv   : dW[tap][co][ci] = sum_pix dY[pix][co] * X[pix*stride + tap_offset][ci]
//   convT  : dW[tap][ci][co] = sum_pix X[pix][ci]  * dYup[2*pix + tap_offset][co]
//
// i.e. D[m][n] (+)= sum_k A[k][m] * B[k][n] with BOTH operands "MN-major": a TMA box
// of an NHWC tensor lands in shared memory as [pixel][channel] rows of 128 bytes
// (SWIZZLE_128B), which is exactly the canonical MN-major SW128 UMMA layout
// (K = pixel rows, 8-row groups 1024 B apart; 128-byte channel blocks LBO apart).
// The M operand walks the un-shifted pixel grid; the N operand is loaded shifted
// (and strided) per tap, re-using one row-slab for the three vertical taps of a
// 3x3 stride-1 filter.  One CTA owns an (m-block of 128) x (n-block) x (tap group)
// output and a strided subset of the pixel tiles (split-K); partial sums are added
// to the fp32 gradient workspace with red.global.add.v4.f32.
//
// Bias gradient: one extra N=16 MMA per K step against a tile of ones.
//
// kTS variant (bf16): the M operand is fed to the MMA from TENSOR MEMORY instead of shared memory.  An M=128 x N=96
// SS-MMA reads 7 KB of operands per 48-clock slot, more than shared memory delivers (measured 56 clk / MMA, tools/probe
// `tsmma`); with A in TMEM only the 3 KB of B come from shared memory and the MMA runs at its 48-clock rate.  The four
// epilogue warps (idle during the mainloop otherwise) transpose each dY tile on the fly: ldmatrix.trans pulls 8x8 blocks
// out of the SW128 tile as (pixel pair, channel) fragments, which is exactly the register layout tcgen05.st.16x256b
// wants for "lane = channel, 32-bit column = two consecutive pixels" (mapping verified by tools/probe `stmap`).  The
// same threads sum the tile for the bias gradient, so the ones-MMA disappears.
#pragma once
#include "vk_common.cuh"
#include "vk_conv_igemm.cuh"   // DTraits

namespace vk {

struct WgradTap {
  int load;      // which N-operand buffer of the stage
  int rowoff;    // first pixel row of this tap's K view inside that buffer
  int tap;       // tap index in the output workspace
};
struct WgradLoad {
  int dx, dy;    // N-operand box origin = tile origin * b_stride + (dx, dy)
};

struct WgradParams {
  int n_img, gh, gw;            // pixel grid walked by the K tiles (conv: output grid; convT: input grid)
  int tiles_x, tiles_y, n_tiles;
  int tw_log2, th;              // K tile = (1<<tw_log2) x th pixels
  int k_rows;                   // pixels per K tile
  int box_rows;                 // pixel rows of one N-operand buffer
  int b_stride;                 // tile origin -> N-operand coordinate multiplier
  int n_a_blocks, n_b_blocks;   // 128-byte channel blocks per operand
  int n_loads, n_taps;          // per tap group
  int n_groups;                 // tap groups (blockIdx.z = group)
  WgradLoad loads[3][3];        // [group][load]
  WgradTap taps[3][9];          // [group][tap] (9: the merged-tap layout has one group with all nine taps)
  int b_kstep_rows;             // N-operand pixel rows between consecutive K steps (tile width, or slab width for halo slabs)
  int shared_tap;               // >= 0: tap slot computed by group g only on K tiles with (tile ^ g) even (split between 2 groups)
  int n_cta;                    // GEMM N per CTA (multiple of 16)
  int n_blocks_n;               // N blocks (blockIdx.y = m_block * n_blocks_n + n_block)
  int acc_stride, tmem_cols;
  int stages, ksplit;
  int m_valid, n_valid;         // valid GEMM rows / cols overall
  int total_taps;
  float* dw;                    // [total_taps][m_valid][n_valid] fp32, accumulated
  float* dbias;                 // [m_valid] fp32 accumulated, or null
  // deterministic split-K: when `partials` is set every K slice (blockIdx.x) stores its partial sums with plain stores
  // into its own [total_taps][m_valid][n_valid] slab (slice_elems apart) instead of red.add into dw; the bias partials
  // go to dbias_partials[(slice * n_groups + group) * m_valid + m].  vk_wgrad_unpack_batched sums the slices in order.
  float* partials;
  float* dbias_partials;
  long long slice_elems;
  int prefetch_dist;            // K tiles of L2 prefetch lookahead (0 = off)
  // merged taps (3x3 stride 1, N operand of at most 16 channels = 32-byte pixel rows, bf16): the N-operand slabs are staged
  // with 32-byte rows (SWIZZLE_32B) and ONE MMA per K step covers three vertical taps — its three 16-column N atoms are
  // the same slab read tw pixel rows apart (LBO = tw * 32 bytes).  The three horizontal shifts are three small slabs of
  // the same stage, so ONE CTA per K slice computes all nine taps and the wide M operand is fetched from L2 once instead
  // of three times (these layers move 100 MB for ~0 FLOPs: the three tap-group CTAs re-reading it was their cost).
  int swapped;                  // operands passed the other way round: write dw transposed ([tap][n][m]), see vk_wgrad_args
  int merge_taps;
  int b_row_bytes;              // bytes per pixel row of an N-operand buffer (128, or 32 with merge_taps)
  int bias_mma;                 // tuning aid (VK_WGRAD_BIAS_MMA=1): bias gradient by the ones-MMA also in bf16
  int debug_skip_epi;           // tuning aid (VK_WGRAD_SKIP_EPI=1): leave the accumulators in TMEM, measure the mainloop alone
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add(float* p, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}
__device__ __forceinline__ void st_v4(float* p, float a, float b, float c, float d) {
  asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ void umma_ts_bf16(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_4d_w(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t saddr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ void tmem_st_16x256b(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float bf16x2_sum(uint32_t v) { return __uint_as_float(v << 16) + __uint_as_float(v & 0xFFFF0000u); }

constexpr int kWgradThreads = 288;   // warps 0, 6-8: TMA producers; 1: MMA; 2-5: epilogue
constexpr int kWgradProducers = 4;

// kRows = pixels per K tile (compile time: the MMA issue loop is fully unrolled with immediate descriptor
// offsets — with a run-time trip count the issuing thread spends ~87 clk per MMA on R2UR / address arithmetic,
// which is what bounded this kernel; tools/probe `mma`)
template <typename DT, int kRows, bool kTS = false>
__global__ void __launch_bounds__(kWgradThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
             const __grid_constant__ WgradParams prm) {
  constexpr bool kTF32 = DTraits<DT>::kTF32;
  static_assert(!kTS || (!kTF32 && kRows % 16 == 0), "the TMEM-A variant is bf16 only");
  constexpr int kACols = kRows / 2;                        // TMEM columns of one transposed A tile (2 pixels per column)
  constexpr int kBlockElems = 128 / int(sizeof(DT));       // channels per 128-byte block
  constexpr int kRowsPerMma = 32 / int(sizeof(DT));        // pixels (K) per UMMA
  constexpr int kAdvance = kRowsPerMma * 128;              // bytes between consecutive K steps
  // MN-major 32-bit operands only exist in the "128B swizzle, 32B atom" layout (K atom = 4 rows)
  constexpr uint32_t kLayout = kTF32 ? 1u : 2u;            // SWIZZLE_128B_BASE32B : SWIZZLE_128B
  constexpr uint32_t kSBO = kTF32 ? 512u : 1024u;          // pitch between K atoms (4 or 8 pixel rows)
  constexpr int kMaxStages = 8;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ __align__(8) uint64_t a_ready[2], a_free[2];    // kTS: transposed A tile written / consumed
  __shared__ uint32_t tmem_base_slot;

  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int group = blockIdx.z;
  const int m_block = blockIdx.y / prm.n_blocks_n;
  const int n_block = blockIdx.y - m_block * prm.n_blocks_n;
  const int m0 = m_block * 128;
  const int n0 = n_block * prm.n_cta;
  // bias gradient (column sums of the M operand): the tap groups of a K slice share it round-robin over the K
  // tiles, so no CTA carries all the extra N=16 MMAs (one group doing them all made it 33 % longer than the rest)
  const bool bias_en = (prm.dbias != nullptr) && n_block == 0;

  const int a_block_bytes = prm.k_rows * 128;
  const int b_block_bytes = prm.box_rows * prm.b_row_bytes;
  const int a_bytes = prm.n_a_blocks * a_block_bytes;
  const int b_bytes = prm.n_b_blocks * b_block_bytes;
  const int stage_bytes = a_bytes + prm.n_loads * b_bytes;
  uint8_t* ones = smem + prm.stages * stage_bytes;         // 2 KB tile of ones (bias gradient)
  const int tiles_per_img = prm.tiles_x * prm.tiles_y;
  // K tiles of this CTA: blockIdx.x, blockIdx.x + ksplit, ...
  const int my_tiles = (prm.n_tiles - int(blockIdx.x) + prm.ksplit - 1) / prm.ksplit;

  // bf16 (non-TS): the bias gradient (column sums of the M operand) is taken by the four otherwise idle epilogue warps
  // straight from the staged tile instead of an extra N=16 MMA per K step against a tile of ones — an N=16 MMA costs the
  // same ~44 clocks of issue time as a wide one, 8 % of the issue time of a 96-channel layer.  Those warps are further
  // consumers of the stage: the "stage free" barrier then expects five arrivals (MMA commit + one per warp).
  constexpr bool kBiasLds = !kTS && !kTF32;
  const bool bias_lds = kBiasLds && bias_en && !prm.bias_mma;
  if (threadIdx.x == 0) {
    for (int s = 0; s < prm.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], bias_lds ? 5 : 1);
    }
    mbar_init(&tmem_full_bar, 1);
    for (int i = 0; i < 2; ++i) mbar_init(&a_ready[i], 4), mbar_init(&a_free[i], 1);
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
  if (warp >= 2 && warp < 6) {
    // ones tile: the value 1.0 in the operand type; any layout of all-ones is all-ones
    const uint32_t one = kTF32 ? 0x3F800000u : 0x3F803F80u;
    uint32_t* o = reinterpret_cast<uint32_t*>(ones);
    for (int i = threadIdx.x - 64; i < 512; i += 128) o[i] = one;
    fence_proxy_async_smem();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_launch_dependents();     // programmatic dependent launch: see conv_v2_kernel
  pdl_wait();

  if (warp == 0 || warp >= 6) {
    // ===================== TMA producers (boxes of a stage spread over 4 warps) =====================
    const int pw = warp == 0 ? 0 : warp - 5;
    if (elect_one()) {
      int op = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int s = it % prm.stages;
        const uint32_t ph = (it / prm.stages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        const int t = blockIdx.x + it * prm.ksplit;
        const int img = t / tiles_per_img;
        const int r = t - img * tiles_per_img;
        const int ty = r / prm.tiles_x, tx = r - ty * prm.tiles_x;
        const int x0 = tx << prm.tw_log2, y0 = ty * prm.th;
        uint8_t* a_s = smem + s * stage_bytes;
        uint8_t* b_s = a_s + a_bytes;
        if (pw == 0) mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
        // L2 prefetch of the tile `prefetch_dist` steps ahead (no shared memory, no barrier): the three (two) CTAs that
        // share a K tile all miss L2 on its first touch, and with 2-3 stages of 70+ KB the ring alone cannot cover a
        // DRAM round trip under load
        if (prm.prefetch_dist > 0 && it + prm.prefetch_dist < my_tiles) {
          const int tp_ = blockIdx.x + (it + prm.prefetch_dist) * prm.ksplit;
          const int pimg = tp_ / tiles_per_img;
          const int pr = tp_ - pimg * tiles_per_img;
          const int pty = pr / prm.tiles_x, ptx = pr - pty * prm.tiles_x;
          const int px0 = ptx << prm.tw_log2, py0 = pty * prm.th;
          int pop = 0;
          for (int j = 0; j < prm.n_a_blocks; ++j, ++pop)
            if (pop % kWgradProducers == pw && group == 0)
              tma_prefetch_l2_4d_w(&tmap_a, m0 + j * kBlockElems, px0, py0, pimg);
          for (int l = 0; l < prm.n_loads; ++l)
            for (int j = 0; j < prm.n_b_blocks; ++j, ++pop)
              if (pop % kWgradProducers == pw)
                tma_prefetch_l2_4d_w(&tmap_b, n0 + j * kBlockElems, px0 * prm.b_stride + prm.loads[group][l].dx,
                                     py0 * prm.b_stride + prm.loads[group][l].dy, pimg);
        }
        for (int j = 0; j < prm.n_a_blocks; ++j, ++op) {
          if (op % kWgradProducers != pw) continue;
          tma_load_4d(a_s + j * a_block_bytes, &tmap_a, &full_bar[s], m0 + j * kBlockElems, x0, y0, img);
        }
        for (int l = 0; l < prm.n_loads; ++l) {
          const WgradLoad& ld = prm.loads[group][l];
          for (int j = 0; j < prm.n_b_blocks; ++j, ++op) {
            if (op % kWgradProducers != pw) continue;
            tma_load_4d(b_s + l * b_bytes + j * b_block_bytes, &tmap_b, &full_bar[s], n0 + j * kBlockElems,
                        x0 * prm.b_stride + ld.dx, y0 * prm.b_stride + ld.dy, img);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // descriptors are advanced by additions on their 32-bit low word only (see vk_conv_igemm.cuh: the
    // uniform datapath in front of tcgen05.mma is slow, keep the dependent chain short)
    {
      const bool leader = elect_one();          // warp-uniform loop, only the tcgen05 instructions are predicated
      const uint32_t idesc = make_idesc(DTraits<DT>::kFmt, 128, prm.n_cta, 1, 1);
      const uint32_t idesc_bias = make_idesc(DTraits<DT>::kFmt, 128, 16, 1, 1);
      const uint64_t desc_hi = make_smem_desc(0, 0, kSBO, kLayout) & 0xFFFFFFFF00000000ull;
      const uint32_t a_lbo = ((uint32_t(a_block_bytes) >> 4) & 0x3FFFu) << 16;
      const bool merged = prm.merge_taps != 0;
      // merged taps: MN-major SWIZZLE_32B atoms (16 channels x 8 pixel rows of 32 bytes), N atoms one tile row apart
      const uint64_t desc_hi_b = merged ? (make_smem_desc(0, 0, 256u, 6u) & 0xFFFFFFFF00000000ull) : desc_hi;
      const uint32_t idesc_merged = make_idesc(DTraits<DT>::kFmt, 128, 48, 1, 1);
      const uint32_t b_lbo = merged ? (((uint32_t(prm.b_kstep_rows) * 32u) >> 4) & 0x3FFFu) << 16
                                    : ((uint32_t(b_block_bytes) >> 4) & 0x3FFFu) << 16;
      const uint32_t ones_lo = ((smem_u32(ones) & 0x3FFFFu) >> 4) | (64u << 16);   // LBO 1024 B
      const uint32_t smem16 = (smem_u32(smem) & 0x3FFFFu) >> 4;
      const uint32_t stage16 = uint32_t(stage_bytes) >> 4, a16 = uint32_t(a_bytes) >> 4;
      constexpr uint32_t kAdv16 = uint32_t(kAdvance) >> 4;
      constexpr int ksteps = kRows / kRowsPerMma;
      const int n_taps = prm.n_taps;
      uint32_t tap_off[5];
#pragma unroll
      for (int tp = 0; tp < 5; ++tp)
        tap_off[tp] = (uint32_t(prm.taps[group][tp].load * b_bytes + prm.taps[group][tp].rowoff * 128) >> 4);
      // B advances by one tile row of the (possibly wider) N-operand buffer per K step; A tiles are dense
      // (b_kstep_rows differs from the tile width only for bf16 halo slabs of 16-pixel-wide tiles: one K step = one slab row)
      const uint32_t b_adv16 = merged ? (uint32_t(kRowsPerMma) * 32u) >> 4
                               : prm.b_kstep_rows == (1 << prm.tw_log2) ? kAdv16 : (uint32_t(prm.b_kstep_rows) * 128u) >> 4;
      const int shared_tap = prm.shared_tap;
      uint32_t shared_accum = 0;
      const uint32_t acc_stride = prm.acc_stride;
      uint32_t accum = 0, bias_accum = 0;
      int s = 0;
      uint32_t ph = 0;
      // kTS: A K-major from TMEM (columns [a_col0 + buf * kACols, ...)), B MN-major from shared memory
      const uint32_t idesc_ts = make_idesc(DTraits<DT>::kFmt, 128, prm.n_cta, 0, 1);
      const uint32_t a_col0 = uint32_t(n_taps) * acc_stride;
      for (int it = 0; it < my_tiles; ++it) {
        mbar_wait(&full_bar[s], ph);
        if constexpr (kTS) mbar_wait(&a_ready[it & 1], uint32_t(it >> 1) & 1u);
        tc_fence_after_sync();
        const uint32_t a_lo = (smem16 + uint32_t(s) * stage16) | a_lbo;
        const uint32_t b_lo = (smem16 + uint32_t(s) * stage16 + a16) | b_lbo;
        if (leader && merged) {
          if constexpr (!kTS) {
            // load l = horizontal shift l - 1: its three vertical taps land in accumulator columns [48 l, 48 l + 48)
            for (int l = 0; l < prm.n_loads; ++l) {
              const uint32_t bd = b_lo + uint32_t(l) * (uint32_t(b_bytes) >> 4);
#pragma unroll
              for (int kk = 0; kk < ksteps; ++kk)
                umma_ss<kTF32>(tmem_base + uint32_t(l) * 48u, desc_hi | (a_lo + kk * kAdv16), desc_hi_b | (bd + kk * b_adv16),
                               idesc_merged, kk == 0 ? accum : 1u);
            }
          }
        }
        if (leader) {
#pragma unroll
          for (int tp = 0; tp < 5; ++tp) {
            if (merged) break;
            const bool is_shared = tp == shared_tap;
            if (tp < n_taps && (!is_shared || (((it ^ group) & 1) == 0))) {
              uint32_t ad = a_lo, bd = b_lo + tap_off[tp];
              uint32_t acc = is_shared ? shared_accum : accum;
              if (is_shared) shared_accum = 1;
#pragma unroll
              for (int kk = 0; kk < ksteps; ++kk) {
                if constexpr (kTS) {
                  umma_ts_bf16(tmem_base + tp * acc_stride, tmem_base + a_col0 + uint32_t(it & 1) * kACols + kk * 8,
                               desc_hi | (bd + kk * b_adv16), idesc_ts, kk == 0 ? acc : 1u);
                } else {
                  umma_ss<kTF32>(tmem_base + tp * acc_stride, desc_hi | (ad + kk * kAdv16), desc_hi | (bd + kk * b_adv16),
                                 idesc, kk == 0 ? acc : 1u);
                }
              }
            }
          }
          if constexpr (kTS) umma_commit(&a_free[it & 1]);
          if (!kTS && (!kBiasLds || prm.bias_mma) && bias_en && (it % prm.n_groups) == group) {
            uint32_t ad = a_lo;
            uint32_t acc = bias_accum;
            bias_accum = 1;
#pragma unroll
            for (int kk = 0; kk < ksteps; ++kk) {
              umma_ss<kTF32>(tmem_base + n_taps * acc_stride, desc_hi | (ad + kk * kAdv16), desc_hi | ones_lo, idesc_bias,
                             kk == 0 ? acc : 1u);
            }
          }
          umma_commit(&empty_bar[s]);
        }
        __syncwarp();
        accum = 1;
        if (++s == prm.stages) s = 0, ph ^= 1;
      }
      if (leader) umma_commit(&tmem_full_bar);
    }
  } else if (warp < 6) {
    // ===================== epilogue: TMEM -> fp32 red.add =====================
    const int q4 = warp & 3;
    const int m = m0 + q4 * 32 + lane;
    float bias_part[4] = {0.f, 0.f, 0.f, 0.f};   // kTS: column sums of dY for channels q4*32 + {0, 8, 16, 24} + lane/4
    if constexpr (kTS) {
      // ---- transposer: dY tile [pixel][channel] (SW128 blocks) -> TMEM [channel lane][pixel-pair column] ----
      // thread L supplies the row address of 8x8 block (L / 8), row (L % 8): blocks 0/1 = pixel rows {0,1,4,5,..} /
      // {2,3,6,7,..} of the first 8 channels, blocks 2/3 the same rows of the next 8 channels, so that after the
      // transposing load register i of thread T is (channel T/4 + 8*(i/2), pixels 4*(T%4) + 2*(i%2) + {0,1}) —
      // the (lane, column) fragment of tcgen05.st.16x256b
      const int blk = lane >> 3, rr8 = lane & 7;
      const int px_in_step = 4 * (rr8 >> 1) + 2 * (blk & 1) + (rr8 & 1);
      const bool lanes_valid = (m0 + q4 * 32) < prm.m_valid;     // a quarter of padding channels has nothing to move
      const uint32_t a_col0 = uint32_t(prm.n_taps) * prm.acc_stride;
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < my_tiles; ++it) {
        mbar_wait(&full_bar[s], ph);
        if (it >= 2) mbar_wait(&a_free[it & 1], uint32_t((it - 2) >> 1) & 1u);
        tc_fence_after_sync();
        if (lanes_valid) {
          const uint32_t a_s = smem_u32(smem + s * stage_bytes);
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int ch = q4 * 32 + hf * 16 + 8 * (blk >> 1);        // first channel of this thread's 8x8 block
            const uint32_t blk_base = a_s + uint32_t(ch >> 6) * uint32_t(a_block_bytes);
            const uint32_t chunk = uint32_t(ch & 63) >> 3;            // 16-byte chunk inside the 128-byte row
#pragma unroll
            for (int kk = 0; kk < kRows / 16; ++kk) {
              const int px = kk * 16 + px_in_step;
              uint32_t r[4];
              ldsm_x4_trans(blk_base + uint32_t(px) * 128u + ((chunk ^ uint32_t(px & 7)) << 4), r);
              tmem_st_16x256b(tmem_base + (uint32_t(q4 * 32 + hf * 16) << 16) + a_col0 + uint32_t(it & 1) * kACols + kk * 8, r);
              bias_part[2 * hf] += bf16x2_sum(r[0]) + bf16x2_sum(r[1]);
              bias_part[2 * hf + 1] += bf16x2_sum(r[2]) + bf16x2_sum(r[3]);
            }
          }
          tmem_st_wait();
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[it & 1]);
        if (++s == prm.stages) s = 0, ph ^= 1;
      }
    }
    // ---- bias gradient from the staged M tiles (bf16): the four epilogue warps, a quarter of the pixel rows each ----
    float bsum[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) bsum[i] = 0.f;
    if (kBiasLds && bias_lds) {
      // lane l owns logical 16-byte chunk (l & 7) = 8 channels of every 128-byte pixel row; this warp walks the rows
      // (l >> 3) + 4 (warp - 2) + 16 i.  SW128: the chunk sits at physical position chunk ^ (bits [7:9] of the row address)
      const uint32_t chunk = uint32_t(lane & 7);
      const int r0 = (lane >> 3) + 4 * (warp - 2);
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < my_tiles; ++it) {
        mbar_wait(&full_bar[s], ph);
        if ((it % prm.n_groups) == group) {
          const uint32_t a_s = smem_u32(smem + s * stage_bytes);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint32_t blk = a_s + uint32_t(j) * uint32_t(a_block_bytes);
#pragma unroll 4
            for (int r = r0; r < kRows; r += 16) {
              const uint32_t row_addr = blk + uint32_t(r) * 128u;
              const uint32_t addr = row_addr + ((chunk ^ ((row_addr >> 7) & 7u)) << 4);
              uint32_t w0, w1, w2, w3;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(addr));
              float* b = bsum + 8 * j;
              b[0] += __uint_as_float(w0 << 16), b[1] += __uint_as_float(w0 & 0xFFFF0000u);
              b[2] += __uint_as_float(w1 << 16), b[3] += __uint_as_float(w1 & 0xFFFF0000u);
              b[4] += __uint_as_float(w2 << 16), b[5] += __uint_as_float(w2 & 0xFFFF0000u);
              b[6] += __uint_as_float(w3 << 16), b[7] += __uint_as_float(w3 & 0xFFFF0000u);
            }
          }
        }
        // the sums above consumed every loaded register, so no load of this warp is still in flight when the stage is
        // handed back to the TMA producers (generic-proxy read -> async-proxy write, see vk_conv_v2.cuh)
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
        if (++s == prm.stages) s = 0, ph ^= 1;
      }
      // rows were dealt to the four 8-lane groups of each warp and to the four warps: add them up in a fixed order
      // (the 2 KB `ones` tile is unused in this variant and serves as the exchange buffer) — lanes 0..7 of warp 2 end
      // with the totals
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float v = bsum[i];
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        bsum[i] = v;
      }
      float* xch = reinterpret_cast<float*>(ones);             // [4 warps][8 lanes][16]
      if (lane < 8) {
#pragma unroll
        for (int i = 0; i < 16; ++i) xch[((warp - 2) * 8 + lane) * 16 + i] = bsum[i];
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (warp == 2 && lane < 8) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          bsum[i] = ((xch[(0 * 8 + lane) * 16 + i] + xch[(1 * 8 + lane) * 16 + i]) + xch[(2 * 8 + lane) * 16 + i]) +
                    xch[(3 * 8 + lane) * 16 + i];
      }
    }
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after_sync();
    const uint32_t lane_addr = tmem_base + (uint32_t(q4 * 32) << 16);
    const bool m_ok = (m < prm.m_valid) && (my_tiles > 0) && !prm.debug_skip_epi;
    const bool vec_ok = (prm.n_valid % 4) == 0;
    // deterministic mode: this K slice's own slab, plain stores (every (tap, m, n) has exactly one writer per slice)
    const bool det = prm.partials != nullptr;
    float* const out_base = det ? prm.partials + static_cast<long long>(blockIdx.x) * prm.slice_elems : prm.dw;
    for (int tp = 0; tp < prm.n_taps; ++tp) {
      // the split tap was accumulated only if this CTA saw a K tile of its parity (tiles it = 0, 1, ...: it ^ group even)
      if (tp == prm.shared_tap && my_tiles <= (group & 1)) continue;
      const int tap = prm.taps[group][tp].tap;
      float* row = out_base + (static_cast<long long>(tap) * prm.m_valid + m) * prm.n_valid;
      for (int jc = 0; jc < prm.n_cta; jc += 16) {
        uint32_t rr[16];
        __syncwarp();
        tmem_ld16(lane_addr + tp * prm.acc_stride + jc, rr);
        tmem_ld_wait();
        const int n = n0 + jc;
        if (prm.swapped) {
          // transposed store: element (tap, n, m); consecutive lanes (m) are consecutive addresses
          if (m_ok) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (n + i < prm.n_valid) {
                float* q = out_base + (static_cast<long long>(tap) * prm.n_valid + (n + i)) * prm.m_valid + m;
                if (det) *q = __uint_as_float(rr[i]);
                else red_add(q, __uint_as_float(rr[i]));
              }
          }
        } else
        if (m_ok && n < prm.n_valid) {
          if (vec_ok && n + 16 <= prm.n_valid) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              if (det)
                st_v4(row + n + i, __uint_as_float(rr[i]), __uint_as_float(rr[i + 1]), __uint_as_float(rr[i + 2]),
                      __uint_as_float(rr[i + 3]));
              else
                red_add_v4(row + n + i, __uint_as_float(rr[i]), __uint_as_float(rr[i + 1]),
                           __uint_as_float(rr[i + 2]), __uint_as_float(rr[i + 3]));
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (n + i < prm.n_valid) {
                if (det) row[n + i] = __uint_as_float(rr[i]);
                else red_add(row + n + i, __uint_as_float(rr[i]));
              }
          }
        }
      }
    }
    if constexpr (kTS) {
      // the four threads that share a channel hold partial sums over different pixel pairs
      if (bias_en && group == 0 && !prm.debug_skip_epi) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v = bias_part[i];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          const int c = m0 + q4 * 32 + (i >> 1) * 16 + (i & 1) * 8 + (lane >> 2);
          if ((lane & 3) == 0 && c < prm.m_valid && my_tiles > 0) red_add(prm.dbias + c, v);
        }
      }
    } else if (kBiasLds && bias_lds) {
      // lanes 0..7 of warp 2 hold the column sums of channels m0 + 64 j + 8 (lane & 7) + i; a group that saw no bias
      // tile holds zeros (deterministic mode: still stored, every slot has exactly one writer)
      if (warp == 2 && lane < 8 && !prm.debug_skip_epi) {
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int c = m0 + 64 * j + 8 * lane + i;
            if (c < prm.m_valid) {
              if (det)
                prm.dbias_partials[(static_cast<long long>(blockIdx.x) * prm.n_groups + group) * prm.m_valid + c] = bsum[8 * j + i];
              else if (group < my_tiles)
                red_add(prm.dbias + c, bsum[8 * j + i]);
            }
          }
      }
    } else if (bias_en && det) {
      // one slot per (K slice, tap group); a group that saw no bias tile (fewer K tiles than groups) stores zero
      float v = 0.f;
      if (group < my_tiles) {
        uint32_t rr[16];
        __syncwarp();
        tmem_ld16(lane_addr + prm.n_taps * prm.acc_stride, rr);
        tmem_ld_wait();
        v = __uint_as_float(rr[0]);
      }
      if (m < prm.m_valid && !prm.debug_skip_epi)
        prm.dbias_partials[(static_cast<long long>(blockIdx.x) * prm.n_groups + group) * prm.m_valid + m] = v;
    } else if (bias_en && group < my_tiles) {
      uint32_t rr[16];
      __syncwarp();
      tmem_ld16(lane_addr + prm.n_taps * prm.acc_stride, rr);
      tmem_ld_wait();
      if (m_ok) red_add(prm.dbias + m, __uint_as_float(rr[0]));
    }
    tc_fence_before_sync();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, prm.tmem_cols);
  }
}

// dW workspace [taps][M][N] -> parameter layout [M][N][taps] (OIHW for conv, [Cin][Cout][kh][kw] for ConvT)
__global__ void wgrad_unpack_kernel(const float* __restrict__ ws, float* __restrict__ out, int taps, int mn,
                                    int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // index over M*N
  if (i >= mn) return;
  for (int t = 0; t < taps; ++t) {
    const float v = ws[static_cast<long long>(t) * mn + i];
    float* o = out + static_cast<long long>(i) * taps + t;
    *o = accumulate ? *o + v : v;
  }
}


// all layers in one launch: blockIdx.y = layer
struct WgradUnpackDesc {
  const float* ws;
  float* out;
  int taps, mn;
  int nslices, pad;         // > 1: ws holds nslices split-K partial slabs, summed here in slice order (deterministic)
  long long slice_stride;   // elements between consecutive slabs
};
// A block transposes 256 consecutive (m, n) elements x taps through shared memory: the workspace rows are read
// coalesced (consecutive elements of one tap), the parameter layout is written as one contiguous run of 256 * taps floats
// (written straight from the per-tap loop it is a 36-byte-stride scatter that costs nine partial-sector passes).
__global__ void __launch_bounds__(256)
wgrad_unpack_batched_kernel(const WgradUnpackDesc* __restrict__ descs, int accumulate) {
  constexpr int kMaxTaps = 9;
  __shared__ float tile[kMaxTaps][257];
  const WgradUnpackDesc d = descs[blockIdx.y];
  auto value = [&](int t, int i) {
    const float* src = d.ws + static_cast<long long>(t) * d.mn + i;
    float v = src[0];
    if (d.nslices > 1) {
      // fixed summation order 0, 1, 2, ...; four independent loads in flight per step
      int s = 1;
      for (; s + 4 <= d.nslices; s += 4) {
        const float x0 = src[(s + 0) * d.slice_stride], x1 = src[(s + 1) * d.slice_stride];
        const float x2 = src[(s + 2) * d.slice_stride], x3 = src[(s + 3) * d.slice_stride];
        v = (((v + x0) + x1) + x2) + x3;
      }
      for (; s < d.nslices; ++s) v += src[s * d.slice_stride];
    }
    return v;
  };
  if (d.taps > kMaxTaps) {                       // not used by this network: plain per-element form
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.mn; i += gridDim.x * blockDim.x)
      for (int t = 0; t < d.taps; ++t) {
        float* o = d.out + static_cast<long long>(i) * d.taps + t;
        const float v = value(t, i);
        *o = accumulate ? *o + v : v;
      }
    return;
  }
  for (int i0 = blockIdx.x * 256; i0 < d.mn; i0 += gridDim.x * 256) {
    const int i = i0 + threadIdx.x;
    if (i < d.mn)
      for (int t = 0; t < d.taps; ++t) tile[t][threadIdx.x] = value(t, i);
    __syncthreads();
    const int cnt = min(256, d.mn - i0) * d.taps;
    float* o = d.out + static_cast<long long>(i0) * d.taps;
    for (int j = threadIdx.x; j < cnt; j += 256) {
      const int ii = j / d.taps, t = j - ii * d.taps;
      const float v = tile[t][ii];
      o[j] = accumulate ? o[j] + v : v;
    }
    __syncthreads();
  }
}

}  // namespace vk
