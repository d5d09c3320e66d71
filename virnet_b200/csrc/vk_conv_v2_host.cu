// vk_conv_v2_host.cu — host side of the persistent convolution kernel (vk_conv_v2.cuh): job /
// stage heuristics under the TMEM (512 columns) and shared-memory (227 KB) budgets, the TMA
// descriptors of the A / B operands and of the epilogue tensors, launch.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "../../include/virnet_b200.h"
#include "vk_conv_v2_launch.h"
#include "vk_host.h"

namespace vk {

namespace {

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
inline int next_pow2_cols(int x) {
  int c = 32;
  while (c < x) c <<= 1;
  return c;
}

int sm_count() {
  static int n = [] {
    int dev = 0, v = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v > 0 ? v : 148;
  }();
  return n;
}

constexpr int kV2SmemBudget = 220 * 1024;   // dynamic smem incl. 1 KB alignment slack (static: ~4.6 KB)

// kernel instantiations live in vk_conv_v2_inst_*.cu (one translation unit per dtype x pair mode, built in parallel)
}  // namespace

// Returns VK_E_UNSUPPORTED when the shape does not fit this kernel (the caller falls back to v1).
int conv_v2_impl(const vk_conv_args* a, void* stream, int phase) {
  const int esize = a->dtype == VK_BF16 ? 2 : 4;
  const int chan_align = 32 / esize;
  if (a->ldx <= 0 || a->ldx % chan_align) return VK_E_BADARG;
  if (a->wrows <= 0 || a->wrows % 16 || a->wrows > 1024) return VK_E_UNSUPPORTED;

  ConvV2Params prm{};
  int taps = 9, us = 1;
  bool slab = false;
  switch (a->kind) {
    case VK_CONV3X3_S1: prm.oh = a->ih, prm.ow = a->iw, prm.a_stride = 1, slab = true; break;
    case VK_CONV3X3_S2: prm.oh = (a->ih + 1) / 2, prm.ow = (a->iw + 1) / 2, prm.a_stride = 2; break;
    case VK_CONVT2X2_S2: prm.oh = a->ih, prm.ow = a->iw, prm.a_stride = 1, taps = 1, us = 2; break;
    case VK_CONV1X1: prm.oh = a->ih, prm.ow = a->iw, prm.a_stride = 1, taps = 1; break;
    case VK_CONV2X2_S2: prm.oh = a->ih / 2, prm.ow = a->iw / 2, prm.a_stride = 2, taps = 4; break;
    case VK_CONV3X3_S2_DGRAD: {
      if (a->out_h <= 0 || a->out_w <= 0 || (a->out_h + 1) / 2 != a->ih || (a->out_w + 1) / 2 != a->iw)
        return VK_E_BADARG;
      if (phase < 0) {
        // merged phases: the tile grid of phase (0, 0), the largest; the other phases' views clip it (TMA bounds)
        if (a->out_h < 2 || a->out_w < 2) return VK_E_UNSUPPORTED;      // an empty phase has no tensor map
        prm.oh = (a->out_h + 1) / 2, prm.ow = (a->out_w + 1) / 2;
        prm.a_stride = 1, us = 2;
        break;
      }
      const int py = phase >> 1, px = phase & 1;
      prm.oh = (a->out_h - py + 1) / 2, prm.ow = (a->out_w - px + 1) / 2;
      prm.a_stride = 1, us = 2;
      if (prm.oh <= 0 || prm.ow <= 0) return 0;
      break;
    }
    default: return VK_E_BADARG;
  }
  const bool is_convT = a->kind == VK_CONVT2X2_S2;
  const bool is_s2d = a->kind == VK_CONV3X3_S2_DGRAD;
  const bool s2d_merged = is_s2d && phase < 0;
  prm.n_img = a->n;
  prm.cout = a->cout;
  prm.wrows = a->wrows;
  prm.cq = is_convT ? a->wrows / 4 : a->wrows;
  if (is_convT && (a->wrows % 64 || prm.cq < a->cout)) return VK_E_BADARG;
  if (!is_convT && a->wrows < a->cout) return VK_E_BADARG;
  const int fine_h = is_s2d ? a->out_h : prm.oh * us;      // spatial dims of the EPI_STD tensors
  const int fine_w = is_s2d ? a->out_w : prm.ow * us;
  if (a->epi == VK_EPI_STD) {
    if (a->ldo <= 0 || a->ldo % (16 / esize) || a->ldo < a->cout) return VK_E_BADARG;
  } else if (a->epi == VK_EPI_NCHW_F32) {
    if (us != 1 || a->out1 == nullptr) return VK_E_BADARG;
  } else {
    return VK_E_BADARG;
  }

  // ---- pixel tile ----
  // 8 x 16 pixel tiles: a tile row is one 8-row UMMA core-matrix group, so tap views of a wider box (halo slab,
  // stride-2 phase slab) are plain row-shifted descriptors with SBO = box width
  const int tw = 8, th = 16;
  prm.tw_log2 = 31 - __builtin_clz(tw);
  prm.th = th;
  prm.tiles_x = (prm.ow + tw - 1) / tw;
  prm.tiles_y = (prm.oh + th - 1) / th;
  prm.n_tiles = prm.tiles_x * prm.tiles_y * a->n;

  // ---- K chunk ----
  const int row_bytes = a->ldx * esize;
  int chunk = 0;
  for (int cb : {128, 64, 32}) {
    if (a->force_chunk_bytes && a->force_chunk_bytes != cb) continue;
    if (row_bytes % cb == 0) { chunk = cb; break; }
  }
  if (chunk == 0) return VK_E_UNSUPPORTED;
  prm.k_chunks = row_bytes / chunk;

  // ---- N split ----
  int n_cta;
  if (is_convT) {
    n_cta = prm.cq;
  } else {
    const int parts = (a->wrows + 255) / 256;
    n_cta = round_up((a->wrows + parts - 1) / parts, 16);
    if (n_cta * parts != a->wrows) {
      n_cta = 0;
      for (int c = 256; c >= 16; c -= 16)
        if (a->wrows % c == 0) { n_cta = c; break; }
    }
  }
  // wide layers whose even split is not a multiple of 32 (288 = 2 x 144) cannot pair; three blocks of 96 can
  if (!is_convT && a->wrows > 256 && n_cta % 32 && a->wrows % 96 == 0 && a->force_impl != 3) n_cta = 96;
  if (n_cta <= 0 || n_cta > 256 || n_cta % 16 || a->wrows % n_cta) return VK_E_UNSUPPORTED;
  if (a->epi == VK_EPI_NCHW_F32 && a->wrows != n_cta) return VK_E_UNSUPPORTED;
  prm.n_cta = n_cta;
  prm.n_blocks = a->wrows / n_cta;
  if (s2d_merged) {
    if (prm.n_blocks != 1) return VK_E_UNSUPPORTED;
    prm.n_blocks = 4, prm.phase_jobs = 1;       // the "N block" index of a job is its phase
  }
  prm.acc_stride = round_up(n_cta, 32);

  // ---- loads / taps ----
  int box_w, box_h;   // A box in pixels (before element strides)
  int nt_opts[3], n_nt = 0;
  if (slab) {
    box_w = tw + 2, box_h = th + 2;
    prm.n_loads = 1;
    prm.loads[0] = {-1, -1, 0};
    prm.a_sbo = box_w * chunk;
    nt_opts[n_nt++] = 9, nt_opts[n_nt++] = 3, nt_opts[n_nt++] = 1;
  } else {
    // full-K mode: an A item is one load (box) with all K chunks, each weight item is one tap with all K chunks
    prm.full_k = 1;
    nt_opts[n_nt++] = 1;
    std::memset(prm.loads, 0, sizeof(prm.loads));
    auto aoff = [&](int dyo, int dxo, int bw) { return uint32_t((dyo * bw + dxo) * chunk) >> 4; };
    if (a->kind == VK_CONV3X3_S2) {
      // the 9 taps read the four (row parity, column parity) phase images of the input: each phase is ONE
      // strided (tw+1) x (th+1) box whose taps are row-shifted views (1, 2, 2 and 4 taps)
      box_w = tw + 1, box_h = th + 1;
      prm.n_loads = 4;
      for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
          ConvV2Load& l = prm.loads[py * 2 + px];
          l.dx = px ? -1 : 0, l.dy = py ? -1 : 0;
          for (int r = 0; r < 3; ++r)
            for (int s2 = 0; s2 < 3; ++s2)
              if ((r != 1) == (py != 0) && (s2 != 1) == (px != 0)) {
                l.tap[l.nb] = r * 3 + s2;
                l.aoff16[l.nb] = aoff(r == 2 ? 1 : 0, s2 == 2 ? 1 : 0, box_w);
                ++l.nb;
              }
        }
    } else if (a->kind == VK_CONV2X2_S2) {
      box_w = tw, box_h = th;
      prm.n_loads = 4;
      for (int t = 0; t < 4; ++t) {
        ConvV2Load& l = prm.loads[t];
        l.dx = t & 1, l.dy = t >> 1, l.nb = 1, l.tap[0] = t, l.aoff16[0] = 0;
      }
    } else if (is_s2d) {
      box_w = tw + 1, box_h = th + 1;
      prm.n_loads = s2d_merged ? 4 : 1;
      for (int ph = 0; ph < 4; ++ph) {
        if (!s2d_merged && ph != phase) continue;
        const int py = ph >> 1, px = ph & 1;
        const int nr = py ? 2 : 1, ns = px ? 2 : 1;
        const int rr[2] = {py ? 0 : 1, 2}, dyv[2] = {py ? 1 : 0, 0};
        const int ss[2] = {px ? 0 : 1, 2}, dxv[2] = {px ? 1 : 0, 0};
        ConvV2Load& l = prm.loads[s2d_merged ? ph : 0];
        for (int i = 0; i < nr; ++i)
          for (int j = 0; j < ns; ++j) {
            l.tap[l.nb] = rr[i] * 3 + ss[j];
            l.aoff16[l.nb] = aoff(dyv[i], dxv[j], box_w);
            ++l.nb;
          }
      }
    } else {
      box_w = tw, box_h = th;
      prm.n_loads = 1;
      prm.loads[0].nb = 1;
    }
    prm.a_sbo = box_w * chunk;
  }
  const int a_rows = box_w * box_h;
  prm.a_tx_bytes = a_rows * chunk;
  prm.a_box_bytes = round_up(prm.a_tx_bytes, 1024);

  // ---- epilogue staging ----
  prm.epi = a->epi;
  prm.has_resid = a->resid != nullptr, prm.has_mask = a->mask != nullptr;
  prm.has_out1 = a->out1 != nullptr, prm.has_out2 = a->out2 != nullptr;
  int ecols = 0, ecb = 0, epi_warp_bytes = 0;
  if (a->epi == VK_EPI_STD) {
    if (us == 2 && prm.has_mask) return VK_E_UNSUPPORTED;
    if ((a->sft_mul != nullptr) != (a->sft_add != nullptr)) return VK_E_BADARG;
    if (a->sft_mul != nullptr && (!prm.has_out2 || prm.has_mask || us != 1 || a->sft_ld < a->cout)) return VK_E_BADARG;
    {  // tensor combinations the epilogue is specialised for (everything the network uses); others -> v1
      const int mode = prm.has_mask | (prm.has_resid << 1) | (prm.has_out1 << 2) | (prm.has_out2 << 3);
      if (mode != 4 && mode != 5 && mode != 6 && mode != 7 && mode != 8 && mode != 12 && mode != 14) return VK_E_UNSUPPORTED;
    }
    ecb = (n_cta * esize) % 64 == 0 ? 64 : 32;
    ecols = ecb / esize;
    if (n_cta % ecols) return VK_E_UNSUPPORTED;
    int off = 0;
    const int unit = 128 * ecb;                 // one 128-pixel tile x ecols channels
    prm.off_r = off; if (prm.has_resid) off += unit;
    prm.off_k = off; if (prm.has_mask) off += unit;
    prm.off_o1 = off; if (prm.has_out1) off += unit;
    prm.off_o2 = off; if (prm.has_out2) off += unit;
    epi_warp_bytes = round_up(std::max(off, unit), 1024);
    prm.ebx = tw, prm.eby = th;
  }
  prm.ecb = ecb, prm.ecols = ecols, prm.n_ech = ecols ? n_cta / ecols : 1;
  prm.epi_warp_bytes = epi_warp_bytes;
  // (epi_bytes depends on the pair decision below)

  // ---- CTA pairs (cta_group::2): M = 256 per MMA, each CTA stages half of the weight rows ----
  // force_impl: 0/2 automatic, 3 single CTAs, 4 pairs
  bool pair = a->force_impl == 4 || (a->force_impl != 3 && n_cta % 32 == 0 && n_cta >= 32);
  if (n_cta % 32) pair = false;                 // each half must be a multiple of 16 rows (N % 16, swizzle atoms)
  const int b_rows = pair ? n_cta / 2 : n_cta;  // weight rows staged per CTA
  const int epi_groups = v2_epi_groups(pair);
  // fast epilogue (specialised bf16 pair slab kernels, vk_conv_v2.cuh): inputs come from global memory into registers,
  // the staging area holds two sets of output buffers; only when a specialised kernel exists for every (chunk, nt)
  // the planner may pick below (nt is restricted accordingly)
  static const bool no_hot_env = std::getenv("VK_V2_NO_HOT") != nullptr;
  static const bool no_fast_env = std::getenv("VK_V2_NO_FAST_EPI") != nullptr;
  const int epi_mode = (prm.has_mask | (prm.has_resid << 1) | (prm.has_out1 << 2) | (prm.has_out2 << 3)) +
                       (a->sft_mul != nullptr ? 16 : 0);
  auto hot_exists = [&](int nt_) {
    return a->dtype == VK_BF16 ? v2_hot_slab_exists(chunk, nt_, epi_mode) : v2_hot_slab_exists_tf32(chunk, nt_, epi_mode);
  };
  const bool fast_epi = pair && a->epi == VK_EPI_STD && ecb == 64 && a->cta_timing == nullptr && !no_hot_env &&
                        !no_fast_env && slab && !prm.has_mask && !prm.has_resid && hot_exists(9) && !a->force_nt;
  prm.fast_epi = fast_epi ? 1 : 0;
  prm.mask_ptr = a->mask;
  prm.ldo_e = a->ldo;
  if (fast_epi) {
    const int unit = 128 * ecb;
    int off = 0;
    prm.off_r = prm.off_k = 0;
    prm.off_o1 = off; if (prm.has_out1) off += unit;
    prm.off_o2 = off; if (prm.has_out2) off += unit;
    epi_warp_bytes = 2 * round_up(std::max(off, unit), 1024);
    prm.epi_warp_bytes = epi_warp_bytes;
  }
  const int epi_bytes = epi_groups * epi_warp_bytes;   // one staging area per epilogue group (4 warps)

  // ---- P, NT and ring depths ----
  const int n_sm = sm_count();
  const int p_max_tmem = std::max(1, 256 / prm.acc_stride);          // double-buffered accumulators
  int p_hi = std::min(2, p_max_tmem);          // the epilogue caches the coordinates of two tiles per job
  if (a->force_tiles_per_cta) p_hi = std::min(p_hi, a->force_tiles_per_cta);
  int best_P = 0, best_nt = 0, best_as = 0, best_bs = 0, best_res = 0, best_kg = 1;
  // full-K mode: the K-chunk group size fixes the accumulation order (load, group, tap, chunk, k), so it must
  // depend on the layer only, never on the batch: the largest group that fits with the TMEM-maximal P, two A
  // slots and three B slots
  int kg_fixed = 1;
  if (!slab) {
    const int p_cap = std::min(2, p_max_tmem);
    const int avail = kV2SmemBudget - 1024 - epi_bytes;
    kg_fixed = 0;
    for (int kg = prm.k_chunks; kg >= 1; --kg) {
      const int a_slot = p_cap * kg * prm.a_box_bytes;
      const int b_slot = round_up(kg * b_rows * chunk, 1024);
      if (2 * a_slot + kV2BProducers * b_slot <= avail) { kg_fixed = kg; break; }
    }
    if (kg_fixed == 0) return VK_E_UNSUPPORTED;
  }
  double best_score = -1.0;
  if (p_hi == 3) p_hi = 2;                      // the epilogue's item <-> tile mapping wants P in {1, 2, 4}
  for (int P = p_hi; P >= 1; P >>= 1) {
    if (a->force_tiles_per_cta && P != p_hi) break;
    const long long groups = (prm.n_tiles + P - 1) / P;
    const long long jobs = (pair ? (groups + 1) / 2 : groups) * prm.n_blocks;
    const long long workers = pair ? n_sm / 2 : n_sm;
    const long long rounds = (jobs + workers - 1) / workers;
    const double eff = double(groups * prm.n_blocks) / double(rounds * workers * (pair ? 2 : 1));
    // slab mode iterates the taps-per-stage options; full-K mode uses the fixed K-chunk group size
    const int n_opts = slab ? n_nt : 1;
    for (int i = 0; i < n_opts; ++i) {
      const int nt = slab ? nt_opts[i] : 1;
      const int kg = slab ? 1 : kg_fixed;
      if (a->force_nt && slab && nt != a->force_nt) continue;
      if (fast_epi && !hot_exists(nt)) continue;        // the staging layout needs the fast kernel
      const int n_sub = kg;                                      // boxes per tile per A item
      const int n_kgroups = slab ? prm.k_chunks : (prm.k_chunks + kg - 1) / kg;
      const int a_slot = P * n_sub * prm.a_box_bytes;
      const int b_slot = round_up((slab ? nt : kg) * b_rows * chunk, 1024);
      int gen_b_items = 0;
      for (int l = 0; l < prm.n_loads; ++l) gen_b_items += prm.loads[l].nb;
      const int total_b_items = slab ? prm.k_chunks * (9 / nt) : gen_b_items * n_kgroups;
      const int total_a_items = slab ? prm.k_chunks : (prm.phase_jobs ? 1 : prm.n_loads) * n_kgroups;
      const int avail = kV2SmemBudget - 1024 - epi_bytes;
      // depth: at least 2 of each; B ring as deep as fits (up to 8), A ring 2 (3 when cheap)
      int as = std::min(2, std::max(1, total_a_items)) ;
      if (as < 2) as = 2;
      int bs = (avail - as * a_slot) / b_slot;
      bs = std::min(bs, kV2MaxStages);
      if (a->force_stages) bs = std::min(bs, a->force_stages);
      // resident weights: one N block and every weight item of a job fits the ring -> loaded once per CTA,
      // no weight traffic (L2 -> smem writes compete with the MMA's operand fetch) after the first job
      const bool resident = prm.n_blocks == 1 && total_b_items <= bs && !a->force_stages &&
                            std::getenv("VK_V2_NO_RESIDENT") == nullptr;
      if (resident) bs = total_b_items;
      // >= kV2BProducers: a B producer warp handles every 3rd item, so with fewer slots it could run two
      // barrier phases ahead of the consumer and its parity wait would alias
      if (!resident && bs < kV2BProducers) continue;
      // spend what is left on a third / fourth A slot
      while (as < 4 && as * a_slot + bs * b_slot + a_slot <= avail && (bs >= 3 || resident)) ++as;
      (void)total_b_items;
      // score: fewer, larger B items (less barrier traffic per MMA), more tiles per job (weight reuse), full waves
      const double mmas_per_b = double(P) * (slab ? nt : kg) * (chunk / 32);
      const double score = eff * (1.0 - 0.12 / P) * (1.0 - 1.5 / (mmas_per_b + 6.0)) * (resident ? 1.15 : 1.0);
      if (score > best_score)
        best_score = score, best_P = P, best_nt = nt, best_as = as, best_bs = bs, best_res = resident ? 1 : 0, best_kg = kg;
    }
  }
  if (best_P == 0) return VK_E_UNSUPPORTED;
  const int P = best_P, nt = best_nt;
  prm.P = P;
  prm.p_log2 = P == 4 ? 2 : (P == 2 ? 1 : 0);
  prm.nb = slab ? 9 / nt : 1;
  prm.kg = best_kg;
  prm.a_slot_bytes = P * best_kg * prm.a_box_bytes;
  prm.b_slot_bytes = round_up((slab ? nt : best_kg) * b_rows * chunk, 1024);
  prm.b_tx_bytes = (slab ? nt : 1) * b_rows * chunk;        // full-K mode: bytes of ONE chunk of a weight item
  prm.a_stages = best_as, prm.b_stages = best_bs;
  prm.b_resident = best_res;
  prm.b_base = prm.a_stages * prm.a_slot_bytes;
  prm.epi_base = prm.b_base + prm.b_stages * prm.b_slot_bytes;
  for (int tap = 0; tap < 9; ++tap)
    prm.a_off16[tap] = slab ? uint32_t(((tap / 3) * box_w + (tap % 3)) * chunk) >> 4 : 0u;
  {
    const int groups = (prm.n_tiles + P - 1) / P;
    prm.n_jobs = (pair ? (groups + 1) / 2 : groups) * prm.n_blocks;
  }
  prm.tmem_cols = next_pow2_cols(2 * P * prm.acc_stride);
  if (prm.tmem_cols > 512) return VK_E_UNSUPPORTED;
  const int smem_bytes = prm.epi_base + epi_bytes + 1024;
  if (smem_bytes > kV2SmemBudget) return VK_E_UNSUPPORTED;

  prm.alpha = a->alpha;
  prm.round_out2 = a->round_out2;
  prm.bias = a->bias;
  prm.sft_mul = a->sft_mul, prm.sft_add = a->sft_add, prm.sft_ld = a->sft_ld;
  prm.out1_ptr = a->out1;
  prm.resid_ptr = a->resid;
  prm.act_expclamp = a->act_expclamp;
  prm.clamp_lo = a->clamp_lo, prm.clamp_hi = a->clamp_hi;
  prm.crop_h = a->crop_h > 0 ? a->crop_h : prm.oh;
  prm.crop_w = a->crop_w > 0 ? a->crop_w : prm.ow;
  prm.timing = a->cta_timing;

  // ---- tensor maps: operands ----
  CUtensorMap ta, tb;
  {
    const uint64_t dims[4] = {uint64_t(a->ldx), uint64_t(a->iw), uint64_t(a->ih), uint64_t(a->n)};
    const uint64_t strides[3] = {uint64_t(row_bytes), uint64_t(row_bytes) * a->iw, uint64_t(row_bytes) * a->iw * a->ih};
    const uint32_t s = prm.a_stride;
    const uint32_t box[4] = {uint32_t(chunk / esize), uint32_t(box_w) * s, uint32_t(box_h) * s, 1u};
    const uint32_t es[4] = {1u, s, s, 1u};
    int r = make_tensor_map(&ta, a->dtype, 4, a->x, dims, strides, box, es, chunk);
    if (r) return r;
  }
  {
    const uint64_t dims[3] = {uint64_t(a->ldx), uint64_t(a->wrows), uint64_t(taps)};
    const uint64_t strides[2] = {uint64_t(row_bytes), uint64_t(row_bytes) * a->wrows};
    const uint32_t box[3] = {uint32_t(chunk / esize), uint32_t(b_rows), uint32_t(nt)};
    const uint32_t es[3] = {1u, 1u, 1u};
    int r = make_tensor_map(&tb, a->dtype, 3, a->w, dims, strides, box, es, chunk);
    if (r) return r;
  }
  // ---- tensor maps: epilogue tensors (one strided view per sub-pixel quadrant when us == 2) ----
  ConvV2Maps em;
  std::memset(&em, 0, sizeof(em));
  if (a->epi == VK_EPI_STD) {
    const uint64_t pix_bytes = uint64_t(a->ldo) * esize;
    auto make_view = [&](CUtensorMap* out, const void* base, int quad) -> int {
      // quadrant (qy, qx) of the fine grid: pixel (us*y + qy, us*x + qx)
      const int qy = us == 2 ? quad >> 1 : 0, qx = us == 2 ? quad & 1 : 0;
      const int vw = us == 2 ? (fine_w - qx + 1) / 2 : fine_w;
      const int vh = us == 2 ? (fine_h - qy + 1) / 2 : fine_h;
      if (vw <= 0 || vh <= 0) return 0;
      const uint8_t* p = reinterpret_cast<const uint8_t*>(base) + (uint64_t(qy) * fine_w + qx) * pix_bytes;
      const uint64_t dims[4] = {uint64_t(a->cout), uint64_t(vw), uint64_t(vh), uint64_t(a->n)};
      const uint64_t strides[3] = {pix_bytes * us, pix_bytes * us * fine_w, pix_bytes * fine_w * fine_h};
      const uint32_t box[4] = {uint32_t(ecols), uint32_t(prm.ebx), uint32_t(prm.eby), 1u};
      const uint32_t es[4] = {1u, 1u, 1u, 1u};
      return make_tensor_map(out, a->dtype, 4, p, dims, strides, box, es, ecb);
    };
    const int q_lo = is_s2d && !s2d_merged ? phase : 0;
    const int nq = is_convT || s2d_merged ? 4 : 1;
    for (int q = 0; q < nq; ++q) {
      int r = 0;
      if (prm.has_out1) r = make_view(&em.out1[q], a->out1, q_lo + q);
      if (!r && prm.has_out2) r = make_view(&em.out2[q], a->out2, q_lo + q);
      if (!r && prm.has_resid) r = make_view(&em.resid[q], a->resid, q_lo + q);
      if (r) return r;
    }
    if (prm.has_mask) {
      int r = make_view(&em.mask, a->mask, 0);
      if (r) return r;
    }
  }

  const int grid = pair ? 2 * int(std::min<long long>(prm.n_jobs, n_sm / 2)) : int(std::min<long long>(prm.n_jobs, n_sm));
  static const bool debug = std::getenv("VK_V2_DEBUG") != nullptr;
  if (debug)
    fprintf(stderr,
            "vk v2: kind=%d n=%d %dx%d ldx=%d wrows=%d | tile %dx%d tiles=%d P=%d n_cta=%d jobs=%d grid=%d | chunk=%d nt=%d "
            "nb=%d a_stages=%d(%d B) b_stages=%d(%d B) epi=%d B/warp ecb=%d tmem=%d smem=%d pair=%d resident=%d fullk=%d mode=%d\n",
            a->kind, a->n, a->ih, a->iw, a->ldx, a->wrows, tw, th, prm.n_tiles, P, n_cta, prm.n_jobs, grid, chunk, nt,
            prm.nb, prm.a_stages, prm.a_slot_bytes, prm.b_stages, prm.b_slot_bytes, epi_warp_bytes, ecb, prm.tmem_cols,
            smem_bytes, int(pair), prm.b_resident, prm.full_k,
            a->epi == VK_EPI_STD ? (prm.has_mask | (prm.has_resid << 1) | (prm.has_out1 << 2) | (prm.has_out2 << 3)) : -1);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // residual-block convolutions (bf16, pairs, slab mode, 64-byte staging rows, one of the four tensor combinations a
  // training step uses): kernels specialised on the epilogue combination; VK_V2_NO_HOT=1 forces the generic kernel
  static const bool no_hot = std::getenv("VK_V2_NO_HOT") != nullptr;
  if (pair && a->epi == VK_EPI_STD && ecb == 64 && prm.timing == nullptr && !no_hot) {
    const int mode = (prm.has_mask | (prm.has_resid << 1) | (prm.has_out1 << 2) | (prm.has_out2 << 3)) +
                     (a->sft_mul != nullptr ? 16 : 0);
    // the specialised slab kernels without input tensors run the double-buffered-staging epilogue: only with its layout
    const bool layout_ok = prm.full_k || ((mode & 3) != 0) || prm.fast_epi;
    if (layout_ok) {
      const int r = a->dtype == VK_BF16 ? v2_launch_bf16_pair_hot(chunk, nt, mode, ta, tb, em, prm, grid, smem_bytes, st)
                                        : v2_launch_tf32_pair_hot(chunk, nt, mode, ta, tb, em, prm, grid, smem_bytes, st);
      if (r != VK_E_UNSUPPORTED) return r;
    }
  }
  if (a->dtype == VK_BF16)
    return pair ? v2_launch_bf16_pair(chunk, nt, ta, tb, em, prm, grid, smem_bytes, st)
                : v2_launch_bf16_single(chunk, nt, ta, tb, em, prm, grid, smem_bytes, st);
  return pair ? v2_launch_tf32_pair(chunk, nt, ta, tb, em, prm, grid, smem_bytes, st)
              : v2_launch_tf32_single(chunk, nt, ta, tb, em, prm, grid, smem_bytes, st);
}

}  // namespace vk
