// vk_wgrad_host.cu — host side of vk_conv_wgrad (tiling, split-K, TMA maps, launch).
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "../../include/virnet_b200.h"
#include "vk_host.h"
#include "vk_wgrad.cuh"

using namespace vk;

namespace {
inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
inline int next_pow2_cols(int x) {
  int c = 32;
  while (c < x) c <<= 1;
  return c;
}
constexpr int kWgradSmemBudget = 216 * 1024;

template <typename DT, int kRows, bool kTS = false>
int launch_wgrad_k(const CUtensorMap& ta, const CUtensorMap& tb, const WgradParams& prm, dim3 grid, int smem_bytes,
                   cudaStream_t st) {
  static int cur = 0;
  auto kern = wgrad_kernel<DT, kRows, kTS>;
  if (smem_bytes > cur) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return int(e);
    cur = smem_bytes;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid, cfg.blockDim = dim3(kWgradThreads), cfg.dynamicSmemBytes = smem_bytes, cfg.stream = st;
  cudaLaunchAttribute attr[2];
  cfg.attrs = attr, cfg.numAttrs = fill_launch_attrs(attr, false, prm.n_tiles <= 48LL * prm.ksplit);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, prm);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return e != cudaSuccess ? int(e) : int(cudaGetLastError());
}
// bf16 with the M operand transposed into tensor memory by the epilogue warps (vk_wgrad.cuh, kTS)
int launch_wgrad_ts(const CUtensorMap& ta, const CUtensorMap& tb, const WgradParams& prm, dim3 grid, int smem_bytes,
                    cudaStream_t st) {
  switch (prm.k_rows) {
    case 128: return launch_wgrad_k<__nv_bfloat16, 128, true>(ta, tb, prm, grid, smem_bytes, st);
    case 64: return launch_wgrad_k<__nv_bfloat16, 64, true>(ta, tb, prm, grid, smem_bytes, st);
    case 32: return launch_wgrad_k<__nv_bfloat16, 32, true>(ta, tb, prm, grid, smem_bytes, st);
    case 16: return launch_wgrad_k<__nv_bfloat16, 16, true>(ta, tb, prm, grid, smem_bytes, st);
    default: return VK_E_UNSUPPORTED;
  }
}
template <typename DT>
int launch_wgrad(const CUtensorMap& ta, const CUtensorMap& tb, const WgradParams& prm, dim3 grid, int smem_bytes,
                 cudaStream_t st) {
  switch (prm.k_rows) {
    case 128: return launch_wgrad_k<DT, 128>(ta, tb, prm, grid, smem_bytes, st);
    case 112: return launch_wgrad_k<DT, 112>(ta, tb, prm, grid, smem_bytes, st);
    case 96: return launch_wgrad_k<DT, 96>(ta, tb, prm, grid, smem_bytes, st);
    case 64: return launch_wgrad_k<DT, 64>(ta, tb, prm, grid, smem_bytes, st);
    case 32: return launch_wgrad_k<DT, 32>(ta, tb, prm, grid, smem_bytes, st);
    case 16: return launch_wgrad_k<DT, 16>(ta, tb, prm, grid, smem_bytes, st);
    default: return VK_E_UNSUPPORTED;
  }
}
}  // namespace

namespace {
// plan_only: stop after the tiling / split-K decisions and report them (vk_conv_wgrad_plan)
int wgrad_impl(const vk_wgrad_args* a, void* stream, bool plan_only, int32_t* slices_out, int32_t* bias_slots_out) {
  if (a == nullptr) return VK_E_BADARG;
  const bool det = a->partials != nullptr || (plan_only && a->max_slices > 0);
  if (!plan_only && (a->a == nullptr || a->b == nullptr || (a->dw == nullptr && a->partials == nullptr))) return VK_E_BADARG;
  if (!plan_only && a->partials != nullptr && (a->max_slices <= 0 || (a->dbias != nullptr && a->dbias_partials == nullptr)))
    return VK_E_BADARG;
  if (a->dtype != VK_BF16 && a->dtype != VK_TF32) return VK_E_BADARG;
  const int esize = a->dtype == VK_BF16 ? 2 : 4;
  const int block_elems = 128 / esize;
  if (a->lda <= 0 || a->ldb <= 0 || (a->lda * esize) % 16 || (a->ldb * esize) % 16) return VK_E_BADARG;
  if (a->m_valid <= 0 || a->n_valid <= 0 || a->m_valid > a->lda || a->n_valid > a->ldb) return VK_E_BADARG;
  if (a->n <= 0 || a->gh <= 0 || a->gw <= 0 || a->bh <= 0 || a->bw <= 0) return VK_E_BADARG;

  WgradParams prm{};
  prm.n_img = a->n;
  prm.gh = a->gh;
  prm.gw = a->gw;
  prm.m_valid = a->m_valid;
  prm.n_valid = a->n_valid;
  prm.dw = a->dw;
  prm.dbias = a->dbias;
  prm.partials = a->partials;
  prm.dbias_partials = a->dbias_partials;
  prm.swapped = a->swapped != 0;
  if (prm.swapped && (a->kind != VK_CONV3X3_S1 || a->dbias != nullptr)) return VK_E_BADARG;
  {
    static const int skip = getenv("VK_WGRAD_SKIP_EPI") != nullptr;
    prm.debug_skip_epi = skip;
    static const int bias_mma = getenv("VK_WGRAD_BIAS_MMA") != nullptr;
    prm.bias_mma = bias_mma;
    static const int pf = getenv("VK_WGRAD_PREFETCH") ? atoi(getenv("VK_WGRAD_PREFETCH")) : 0;   // measured: slower
    prm.prefetch_dist = pf;
  }

  bool slab = false;
  prm.shared_tap = -1;
  // 3x3 stride 1: ONE halo slab of the N operand ((tw+2) x (th+2) pixels) serves all nine taps through descriptors whose
  // start is shifted by r * (tw+2) + s pixel rows (the swizzle is a function of the absolute address also for MN-major
  // operands: tools/probe `tsmma`, shifts 1..19).  Two tap groups of 4 taps + the centre tap, which the two CTAs of a
  // K slice compute on alternating K tiles: each operand byte is fetched from L2 by 2 CTAs instead of 3 and every CTA
  // issues 4.5 taps per tile.  Opt-in with VK_WGRAD_SLAB9=1 (default: three tap groups of three vertical taps, one slab per horizontal shift).
  static const bool slab3 = getenv("VK_WGRAD_SLAB9") == nullptr;   // measured: no faster (profiles/r02_wgrad_experiments.txt): opt-in
  bool slab9 = false;
  switch (a->kind) {
    case VK_CONV3X3_S1:
      slab = true;
      prm.b_stride = 1, prm.n_groups = 3, prm.n_loads = 1, prm.n_taps = 3, prm.total_taps = 9;
      if (!slab3 && a->dtype == VK_BF16 && !det && !a->swapped) slab9 = true, prm.n_groups = 2, prm.n_taps = 5, prm.shared_tap = 4;
      break;
    case VK_CONV3X3_S2:
      prm.b_stride = 2, prm.n_groups = 3, prm.n_loads = 3, prm.n_taps = 3, prm.total_taps = 9;
      for (int r = 0; r < 3; ++r)
        for (int s = 0; s < 3; ++s) {
          prm.loads[r][s].dx = s - 1, prm.loads[r][s].dy = r - 1;
          prm.taps[r][s].load = s, prm.taps[r][s].rowoff = 0, prm.taps[r][s].tap = r * 3 + s;
        }
      break;
    case VK_CONVT2X2_S2:
      prm.b_stride = 2, prm.n_groups = 2, prm.n_loads = 2, prm.n_taps = 2, prm.total_taps = 4;
      for (int dy = 0; dy < 2; ++dy)
        for (int dx = 0; dx < 2; ++dx) {
          prm.loads[dy][dx].dx = dx, prm.loads[dy][dx].dy = dy;
          prm.taps[dy][dx].load = dx, prm.taps[dy][dx].rowoff = 0, prm.taps[dy][dx].tap = dy * 2 + dx;
        }
      break;
    case VK_CONV1X1:
      prm.b_stride = 1, prm.n_groups = 1, prm.n_loads = 1, prm.n_taps = 1, prm.total_taps = 1;
      prm.loads[0][0].dx = 0, prm.loads[0][0].dy = 0;
      prm.taps[0][0].load = 0, prm.taps[0][0].rowoff = 0, prm.taps[0][0].tap = 0;
      break;
    default: return VK_E_BADARG;
  }

  // ---- N split under the TMEM budget: n_taps * round32(n_cta) + 32 (bias) <= 512; the TMEM-A variant (bf16) keeps
  // two transposed M tiles of k_rows / 2 columns there instead of the bias accumulator ----
  // (the TMEM-A variant leaves room for 4 accumulators of 96 columns only: it stays with the three-group layout)
  static const bool no_ts = getenv("VK_WGRAD_TS") == nullptr;          // opt-in: same speed on the big layers (power-bound)
  const bool ts = a->dtype == VK_BF16 && !no_ts && !slab9 && !det;
  const int n_pad = round_up(a->n_valid, 16);
  const int max_n = std::min(256, ((512 - (ts ? 128 : 32)) / prm.n_taps) / 32 * 32);
  const int parts = (n_pad + max_n - 1) / max_n;
  const int n_cta = round_up((n_pad + parts - 1) / parts, 16);
  prm.n_cta = n_cta;
  prm.n_blocks_n = parts;
  prm.acc_stride = round_up(n_cta, 32);
  prm.tmem_cols = next_pow2_cols(prm.n_taps * prm.acc_stride + (ts ? 128 : 32));
  if (prm.tmem_cols > 512) return VK_E_UNSUPPORTED;
  prm.n_a_blocks = 128 / block_elems;
  prm.n_b_blocks = (n_cta + block_elems - 1) / block_elems;
  // merged vertical taps for narrow N operands (head, first SNet layer): see WgradParams::merge_taps
  static const bool no_merge = getenv("VK_WGRAD_NO_MERGE") != nullptr;
  const bool merge = a->kind == VK_CONV3X3_S1 && a->dtype == VK_BF16 && !slab9 && !ts && !no_merge && n_cta == 16 &&
                     a->ldb == 16 && parts == 1;
  prm.merge_taps = merge ? 1 : 0;
  prm.b_row_bytes = merge ? 32 : 128;
  if (merge) {
    prm.n_groups = 1, prm.n_loads = 3, prm.n_taps = 9;    // one CTA per K slice: three small slabs, nine taps
    prm.acc_stride = 16;                                  // tap (r, s) = 16-column atom 3 s + r
    prm.tmem_cols = next_pow2_cols(9 * 16 + 32);
  }
  const int m_blocks = (a->m_valid + 127) / 128;

  // ---- K tile and stages under the smem budget ----
  // (tw, th): 128, 112, 96, 64, 32, 16 pixels; the 7- and 6-row tiles exist for the 9-tap halo slab only (three ring
  // stages of A tile + slab fit the shared memory with them, two with 8 rows)
  static const int cand[6][2] = {{16, 8}, {16, 7}, {16, 6}, {16, 4}, {8, 4}, {8, 2}};
  int tw = 0, th = 0, stages = 0;
  for (int want = 3; want >= 1 && !stages; --want) {
    for (int i = 0; i < 6 && !stages; ++i) {
      const int k_rows = cand[i][0] * cand[i][1];
      if (a->force_k_rows && a->force_k_rows != k_rows) continue;
      if ((cand[i][1] == 7 || cand[i][1] == 6) && !slab9) continue;
      if (k_rows * esize < 32 * 8 / 8 * 8) { /* at least one UMMA K step */ }
      const int box_rows = slab9 ? (cand[i][1] + 2) * (cand[i][0] + 2) : slab ? (cand[i][1] + 2) * cand[i][0] : k_rows;
      const int stage = prm.n_a_blocks * k_rows * 128 + prm.n_loads * prm.n_b_blocks * box_rows * prm.b_row_bytes;
      int st = std::min(8, kWgradSmemBudget / stage);
      if (a->force_stages) st = std::min(st, a->force_stages);
      if (slab9 && cand[i][0] != 16) continue;           // a K step (16 pixels) must be one contiguous slab row
      if (st >= want) tw = cand[i][0], th = cand[i][1], stages = st;
    }
  }
  if (!stages) return VK_E_UNSUPPORTED;
  prm.tw_log2 = 31 - __builtin_clz(tw);
  prm.th = th;
  prm.k_rows = tw * th;
  prm.box_rows = slab9 ? (th + 2) * (tw + 2) : slab ? (th + 2) * tw : tw * th;
  prm.b_kstep_rows = slab9 ? tw + 2 : tw;
  prm.stages = stages;
  if (slab9) {
    // group 0: taps 0..3 + centre (4) on even K tiles; group 1: taps 5..8 + centre on odd K tiles
    for (int g = 0; g < 2; ++g) {
      prm.loads[g][0].dx = -1, prm.loads[g][0].dy = -1;
      for (int i = 0; i < 5; ++i) {
        const int tap = i == 4 ? 4 : (g == 0 ? i : 5 + i);
        prm.taps[g][i].load = 0, prm.taps[g][i].rowoff = (tap / 3) * (tw + 2) + (tap % 3), prm.taps[g][i].tap = tap;
      }
    }
  } else if (slab && merge) {
    for (int l = 0; l < 3; ++l) {
      prm.loads[0][l].dx = l - 1, prm.loads[0][l].dy = -1;
      for (int r = 0; r < 3; ++r) {
        WgradTap& t = prm.taps[0][l * 3 + r];
        t.load = l, t.rowoff = r * tw, t.tap = a->swapped ? 8 - (r * 3 + l) : r * 3 + l;
      }
    }
  } else if (slab) {
    for (int s = 0; s < 3; ++s) {
      prm.loads[s][0].dx = s - 1, prm.loads[s][0].dy = -1;
      for (int r = 0; r < 3; ++r)
        prm.taps[s][r].load = 0, prm.taps[s][r].rowoff = r * tw, prm.taps[s][r].tap = a->swapped ? 8 - (r * 3 + s) : r * 3 + s;
    }
  }
  prm.tiles_x = (a->gw + tw - 1) / tw;
  prm.tiles_y = (a->gh + th - 1) / th;
  prm.n_tiles = prm.tiles_x * prm.tiles_y * a->n;

  const int base_ctas = m_blocks * parts * prm.n_groups;
  // one CTA per SM (its smem ring fills the SM): never launch a partial second wave
  int n_sm = 148;
  {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (n_sm <= 0) n_sm = 148;
  }
  int ksplit = std::max(1, n_sm / base_ctas);
  if (a->force_ksplit) ksplit = a->force_ksplit;
  ksplit = std::min(ksplit, prm.n_tiles);
  if (det) ksplit = std::min(ksplit, a->max_slices);   // one partial slab per K slice
  prm.ksplit = ksplit;
  prm.slice_elems = static_cast<long long>(prm.total_taps) * a->m_valid * a->n_valid;
  if (slices_out != nullptr) *slices_out = ksplit;
  if (bias_slots_out != nullptr) *bias_slots_out = ksplit * prm.n_groups;
  if (plan_only) return 0;

  // ---- tensor maps: 128-byte channel blocks; SWIZZLE_128B (bf16) / 128B_ATOM_32B (fp32, code 129) ----
  const int sw_code = a->dtype == VK_BF16 ? 128 : 129;
  CUtensorMap ta, tb;
  {
    const uint64_t rb = uint64_t(a->lda) * esize;
    const uint64_t dims[4] = {uint64_t(a->lda), uint64_t(a->gw), uint64_t(a->gh), uint64_t(a->n)};
    const uint64_t strides[3] = {rb, rb * a->gw, rb * a->gw * a->gh};
    const uint32_t box[4] = {uint32_t(block_elems), uint32_t(tw), uint32_t(th), 1u};
    const uint32_t es[4] = {1u, 1u, 1u, 1u};
    int r = make_tensor_map(&ta, a->dtype, 4, a->a, dims, strides, box, es, sw_code);
    if (r) return r;
  }
  {
    const uint64_t rb = uint64_t(a->ldb) * esize;
    const uint64_t dims[4] = {uint64_t(a->ldb), uint64_t(a->bw), uint64_t(a->bh), uint64_t(a->n)};
    const uint64_t strides[3] = {rb, rb * a->bw, rb * a->bw * a->bh};
    const uint32_t s = prm.b_stride;
    const uint32_t box_h = slab ? th + 2 : th;
    const uint32_t box_w = slab9 ? tw + 2 : tw;
    const uint32_t box[4] = {uint32_t(merge ? 16 : block_elems), box_w * s, box_h * s, 1u};
    const uint32_t es[4] = {1u, s, s, 1u};
    int r = make_tensor_map(&tb, a->dtype, 4, a->b, dims, strides, box, es, merge ? 32 : sw_code);
    if (r) return r;
  }

  const int stage_bytes = prm.n_a_blocks * prm.k_rows * 128 + prm.n_loads * prm.n_b_blocks * prm.box_rows * prm.b_row_bytes;
  const int smem_bytes = stages * stage_bytes + 2048 + 1024;
  dim3 grid(ksplit, m_blocks * parts, prm.n_groups);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (ts) return launch_wgrad_ts(ta, tb, prm, grid, smem_bytes, st);
  if (a->dtype == VK_BF16) return launch_wgrad<__nv_bfloat16>(ta, tb, prm, grid, smem_bytes, st);
  return launch_wgrad<float>(ta, tb, prm, grid, smem_bytes, st);
}
}  // namespace

extern "C" int vk_conv_wgrad(const vk_wgrad_args* a, void* stream) { return wgrad_impl(a, stream, false, nullptr, nullptr); }

extern "C" int vk_conv_wgrad_plan(const vk_wgrad_args* a, int32_t* slices, int32_t* bias_slots) {
  if (slices == nullptr || bias_slots == nullptr) return VK_E_BADARG;
  return wgrad_impl(a, nullptr, true, slices, bias_slots);
}

extern "C" int vk_wgrad_unpack(const float* ws, float* out, int32_t taps, int32_t m, int32_t n, int32_t accumulate,
                               void* stream) {
  if (ws == nullptr || out == nullptr || taps <= 0 || m <= 0 || n <= 0) return VK_E_BADARG;
  const int mn = m * n;
  wgrad_unpack_kernel<<<(mn + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(ws, out, taps, mn,
                                                                                            accumulate);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return int(cudaGetLastError());
}

extern "C" int vk_wgrad_unpack_batched(const void* descs_dev, int32_t ndesc, int64_t max_mn, int32_t accumulate,
                                       void* stream) {
  if (descs_dev == nullptr || ndesc <= 0 || max_mn <= 0) return VK_E_BADARG;
  static_assert(sizeof(WgradUnpackDesc) == sizeof(vk_unpack_desc), "descriptor layout");
  // blocks per layer: enough for ~16 blocks per SM over the whole table (a bucket of a few large layers gets more
  // blocks per layer than the full 82-layer table)
  const int64_t cap = std::max<int64_t>(64, (148 * 16) / ndesc);
  const int bx = int(std::min<int64_t>((max_mn + 255) / 256, cap));
  wgrad_unpack_batched_kernel<<<dim3(bx, ndesc), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const WgradUnpackDesc*>(descs_dev), accumulate);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return int(cudaGetLastError());
}

extern "C" uint32_t vk_sizeof_wgrad_args(void) { return uint32_t(sizeof(vk_wgrad_args)); }
