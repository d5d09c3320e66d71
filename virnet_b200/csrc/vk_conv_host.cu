// vk_conv_host.cu — host side of vk_conv_igemm: tiling heuristics, TMA
// descriptor cache, launch.  See include/virnet_b200.h for the contract.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "../../include/virnet_b200.h"
#include "vk_conv_igemm.cuh"
#include "vk_host.h"

namespace vk {

std::atomic<uint64_t> g_launch_count{0};

// ---------------------------------------------------------------------------
// cuTensorMapEncodeTiled through the runtime (libcuda is not linked)
// ---------------------------------------------------------------------------
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode_fn() {
  static EncodeFn fn = []() -> EncodeFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<EncodeFn>(p);
  }();
  return fn;
}

struct MapKey {
  uint64_t v[16];
  bool operator==(const MapKey& o) const { return std::memcmp(v, o.v, sizeof(v)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (uint64_t x : k.v) {
      h ^= x;
      h *= 1099511628211ull;
    }
    return static_cast<size_t>(h);
  }
};
static std::mutex g_map_mu;
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;

int make_tensor_map(CUtensorMap* out, int dtype, int rank, const void* ptr, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estrides,
                    int swizzle_bytes) {
  MapKey key{};
  key.v[0] = reinterpret_cast<uint64_t>(ptr);
  key.v[1] = (uint64_t(dtype) << 32) | (uint64_t(rank) << 16) | uint64_t(swizzle_bytes);
  for (int i = 0; i < rank; ++i) {
    key.v[2 + i] = dims[i];
    key.v[7 + i] = (uint64_t(box[i]) << 32) | estrides[i];
    if (i > 0) key.v[11 + i] = strides_bytes[i - 1];
  }
  {
    std::lock_guard<std::mutex> g(g_map_mu);
    auto it = g_map_cache.find(key);
    if (it != g_map_cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeFn enc = get_encode_fn();
  if (enc == nullptr) return VK_E_NODRIVER;
  CUtensorMapSwizzle sw = swizzle_bytes == 129  ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                          : swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUtensorMapDataType dt = dtype == VK_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) gd[i] = dims[i], bx[i] = box[i], es[i] = estrides[i];
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUtensorMap m;
  CUresult r = enc(&m, dt, rank, const_cast<void*>(ptr), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr,
            "vk: cuTensorMapEncodeTiled failed (%d): rank=%d dims=[%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u] "
            "es=[%u,%u,%u,%u] sw=%d ptr=%p\n",
            int(r), rank, (unsigned long long)gd[0], (unsigned long long)gd[1],
            (unsigned long long)(rank > 2 ? gd[2] : 0), (unsigned long long)(rank > 3 ? gd[3] : 0), bx[0], bx[1],
            rank > 2 ? bx[2] : 0, rank > 3 ? bx[3] : 0, es[0], es[1], rank > 2 ? es[2] : 0, rank > 3 ? es[3] : 0,
            swizzle_bytes, ptr);
    return VK_E_BADARG;
  }
  {
    std::lock_guard<std::mutex> g(g_map_mu);
    if (g_map_cache.size() > 65536) g_map_cache.clear();
    g_map_cache.emplace(key, m);
  }
  *out = m;
  return 0;
}

// ---------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------
template <typename K>
static int ensure_smem(K kernel, int bytes) {
  static std::mutex mu;
  static std::unordered_map<const void*, int> cur;
  std::lock_guard<std::mutex> g(mu);
  const void* key = reinterpret_cast<const void*>(kernel);
  auto it = cur.find(key);
  if (it != cur.end() && it->second >= bytes) return 0;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return int(e);
  cur[key] = bytes;
  return 0;
}

template <typename DT, int kChunk>
static int launch_conv(const CUtensorMap& ta, const CUtensorMap& tb, const ConvIgemmParams& prm, dim3 grid,
                       int smem_bytes, cudaStream_t st) {
  auto kern = conv_igemm_kernel<DT, kChunk>;
  int r = ensure_smem(kern, smem_bytes);
  if (r) return r;
  kern<<<grid, kConvThreads, smem_bytes, st>>>(ta, tb, prm);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return int(cudaGetLastError());
}

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline int next_pow2_cols(int x) {
  int c = 32;
  while (c < x) c <<= 1;
  return c;
}

constexpr int kSmemBudget = 200 * 1024;   // dynamic smem we allow one CTA (<= 227 KB - static)

}  // namespace vk

using namespace vk;

static int conv_igemm_impl(const vk_conv_args* a, void* stream, int phase);
namespace vk {
int conv_v2_impl(const vk_conv_args* a, void* stream, int phase);   // vk_conv_v2_host.cu
}

extern "C" int vk_conv_igemm(const vk_conv_args* a, void* stream) {
  if (a == nullptr || a->x == nullptr || a->w == nullptr) return VK_E_BADARG;
  if (a->kind == VK_CONV3X3_S2_DGRAD) {
    // all four (row, column) parity phases of the fine grid as ONE persistent launch (a job = tile group x phase): dY is
    // fetched from HBM once instead of four times; shapes it declines (an empty phase, several N blocks) run per phase
    if (a->force_impl != 1 && std::getenv("VK_S2D_PER_PHASE") == nullptr) {
      if (a->dtype != VK_BF16 && a->dtype != VK_TF32) return VK_E_BADARG;
      if (a->n <= 0 || a->ih <= 0 || a->iw <= 0 || a->cout <= 0) return VK_E_BADARG;
      const int r = conv_v2_impl(a, stream, -1);
      if (r != VK_E_UNSUPPORTED) return r;
    }
    for (int phase = 0; phase < 4; ++phase) {
      int r = conv_igemm_impl(a, stream, phase);
      if (r) return r;
    }
    return 0;
  }
  return conv_igemm_impl(a, stream, 0);
}

static int conv_igemm_impl(const vk_conv_args* a, void* stream, int phase) {
  if (a->dtype != VK_BF16 && a->dtype != VK_TF32) return VK_E_BADARG;
  const int esize = a->dtype == VK_BF16 ? 2 : 4;
  const int chan_align = 32 / esize;                       // one UMMA K step = 32 bytes
  if (a->ldx <= 0 || a->ldx % chan_align) return VK_E_BADARG;
  if (a->wrows <= 0 || a->wrows % 16) return VK_E_BADARG;
  if (a->n <= 0 || a->ih <= 0 || a->iw <= 0 || a->cout <= 0) return VK_E_BADARG;
  // persistent kernel first (force_impl: 0 auto, 1 = v1 only, 2 = v2 only); v1 serves what v2 declines
  if (a->sft_mul != nullptr && a->epi != VK_EPI_STD) return VK_E_BADARG;
  if (a->force_impl != 1) {
    const int r = conv_v2_impl(a, stream, phase);
    if (r != VK_E_UNSUPPORTED || a->force_impl >= 2 || a->sft_mul != nullptr) return r;   // v1 has no SFT epilogue
  }

  ConvIgemmParams prm{};
  int taps = 9;
  int us = 1;
  switch (a->kind) {
    case VK_CONV3X3_S1: prm.oh = a->ih, prm.ow = a->iw, prm.a_stride = 1; break;
    case VK_CONV3X3_S2: prm.oh = (a->ih + 1) / 2, prm.ow = (a->iw + 1) / 2, prm.a_stride = 2; break;
    case VK_CONVT2X2_S2: prm.oh = a->ih, prm.ow = a->iw, prm.a_stride = 1, taps = 1, us = 2; break;
    case VK_CONV1X1: prm.oh = a->ih, prm.ow = a->iw, prm.a_stride = 1, taps = 1; break;
    case VK_CONV2X2_S2: prm.oh = a->ih / 2, prm.ow = a->iw / 2, prm.a_stride = 2, taps = 4; break;
    case VK_CONV3X3_S2_DGRAD: {
      // phase (py, px): fine pixels y = 2u + py, x = 2v + px; the M tiles walk (u, v)
      if (a->out_h <= 0 || a->out_w <= 0 || (a->out_h + 1) / 2 != a->ih || (a->out_w + 1) / 2 != a->iw)
        return VK_E_BADARG;
      const int py = phase >> 1, px = phase & 1;
      prm.oh = (a->out_h - py + 1) / 2, prm.ow = (a->out_w - px + 1) / 2;
      prm.a_stride = 1, us = 2;
      if (prm.oh <= 0 || prm.ow <= 0) return 0;   // empty phase (1-pixel-wide images)
      break;
    }
    default: return VK_E_BADARG;
  }
  prm.n_img = a->n;
  prm.us = us;
  prm.cout = a->cout;
  const bool is_convT = a->kind == VK_CONVT2X2_S2;
  prm.cq = is_convT ? a->wrows / 4 : a->wrows;
  prm.quad_base = a->kind == VK_CONV3X3_S2_DGRAD ? phase : 0;
  prm.out_h = a->kind == VK_CONV3X3_S2_DGRAD ? a->out_h : prm.oh * us;
  prm.out_w = a->kind == VK_CONV3X3_S2_DGRAD ? a->out_w : prm.ow * us;
  if (is_convT && (a->wrows % 64 || prm.cq < a->cout)) return VK_E_BADARG;
  if (!is_convT && a->wrows < a->cout) return VK_E_BADARG;
  if (a->kind == VK_CONV3X3_S2_DGRAD || a->kind == VK_CONV2X2_S2) taps = a->kind == VK_CONV2X2_S2 ? 4 : 9;
  if (a->epi == VK_EPI_STD) {
    if (a->ldo <= 0 || a->ldo % (16 / esize)) return VK_E_BADARG;   // 16-byte vector stores
    if (a->ldo < a->cout) return VK_E_BADARG;
  } else if (a->epi == VK_EPI_NCHW_F32) {
    if (us != 1 || a->out1 == nullptr) return VK_E_BADARG;
    prm.out_h = prm.oh, prm.out_w = prm.ow;
  } else {
    return VK_E_BADARG;
  }

  // ---- pixel tile shape: minimise padded pixels, prefer square-ish ----
  static const int cand_tw[5] = {16, 8, 32, 64, 128};
  int best_tw = 16;
  long long best_cost = -1;
  for (int i = 0; i < 5; ++i) {
    const int tw = cand_tw[i], th = 128 / tw;
    if (a->force_tw && a->force_tw != tw) continue;
    if (prm.a_stride == 2 && tw * 2 > 256) continue;                // TMA box limit with elementStrides
    const long long tiles = (long long)((prm.ow + tw - 1) / tw) * ((prm.oh + th - 1) / th);
    // halo overhead of the slab loads: (th+2)/th rows fetched per tile row
    const long long cost = tiles * (a->kind == VK_CONV3X3_S1 ? (th + 2) * tw : th * tw);
    if (best_cost < 0 || cost < best_cost) best_cost = cost, best_tw = tw;
  }
  const int tw = best_tw, th = 128 / tw;
  prm.tw_log2 = 31 - __builtin_clz(tw);
  prm.th = th;
  prm.tiles_x = (prm.ow + tw - 1) / tw;
  prm.tiles_y = (prm.oh + th - 1) / th;
  prm.n_tiles = prm.tiles_x * prm.tiles_y * a->n;

  // ---- loads ----
  int box_h = th;
  if (a->kind == VK_CONV3X3_S1) {
    box_h = th + 2;
    prm.n_loads = 3;
    prm.b_taps = 3;
    for (int s = 0; s < 3; ++s) {
      ConvLoad& l = prm.loads[s];
      l.dx = s - 1, l.dy = -1, l.ntaps = 3;
      for (int r = 0; r < 3; ++r) l.tap[r] = r * 3 + s, l.rowoff[r] = r * tw;
    }
  } else if (a->kind == VK_CONV3X3_S2) {
    prm.n_loads = 9;
    prm.b_taps = 1;
    for (int r = 0; r < 3; ++r)
      for (int s = 0; s < 3; ++s) {
        ConvLoad& l = prm.loads[r * 3 + s];
        l.dx = s - 1, l.dy = r - 1, l.ntaps = 1, l.tap[0] = r * 3 + s, l.rowoff[0] = 0;
      }
  } else if (a->kind == VK_CONV2X2_S2) {
    prm.n_loads = 4;
    prm.b_taps = 1;
    for (int t = 0; t < 4; ++t) {
      ConvLoad& l = prm.loads[t];
      l.dx = t & 1, l.dy = t >> 1, l.ntaps = 1, l.tap[0] = t, l.rowoff[0] = 0;
    }
  } else if (a->kind == VK_CONV3X3_S2_DGRAD) {
    // fine row y = 2*oy - 1 + r: even rows see r = 1 (oy = u); odd rows r = 0 (oy = u + 1) and r = 2 (oy = u)
    const int py = phase >> 1, px = phase & 1;
    const int nr = py ? 2 : 1, ns = px ? 2 : 1;
    const int rr[2] = {py ? 0 : 1, 2}, dyv[2] = {py ? 1 : 0, 0};
    const int ss[2] = {px ? 0 : 1, 2}, dxv[2] = {px ? 1 : 0, 0};
    prm.n_loads = 0;
    prm.b_taps = 1;
    for (int i = 0; i < nr; ++i)
      for (int j = 0; j < ns; ++j) {
        ConvLoad& l = prm.loads[prm.n_loads++];
        l.dx = dxv[j], l.dy = dyv[i], l.ntaps = 1, l.tap[0] = rr[i] * 3 + ss[j], l.rowoff[0] = 0;
      }
  } else {
    prm.n_loads = 1;
    prm.b_taps = 1;
    ConvLoad& l = prm.loads[0];
    l.dx = 0, l.dy = 0, l.ntaps = 1, l.tap[0] = 0, l.rowoff[0] = 0;
  }
  prm.box_rows = box_h * tw;

  // ---- N split ----
  int n_cta;
  if (is_convT) {
    n_cta = prm.cq <= 256 ? prm.cq : 0;
    if (n_cta == 0) return VK_E_UNSUPPORTED;
  } else {
    const int parts = (a->wrows + 255) / 256;
    n_cta = round_up((a->wrows + parts - 1) / parts, 16);
    if (n_cta * parts != a->wrows) {
      // fall back to the largest multiple of 16 <= 256 that divides wrows
      n_cta = 0;
      for (int c = 256; c >= 16; c -= 16)
        if (a->wrows % c == 0) { n_cta = c; break; }
    }
  }
  if (n_cta <= 0 || n_cta > 256 || n_cta % 16 || a->wrows % n_cta) return VK_E_UNSUPPORTED;
  prm.n_cta = n_cta;
  prm.acc_stride = round_up(n_cta, 32);
  const int n_blocks = a->wrows / n_cta;

  // ---- P (tiles per CTA), K chunk and stages under the smem / TMEM budget ----
  const int row_bytes = a->ldx * esize;
  int chunk = 0, P = 0, stages = 0;
  {
    const int max_p_tmem = std::max(1, 512 / prm.acc_stride);
    int p_hi = std::min(max_p_tmem, 4);
    if (a->force_tiles_per_cta) p_hi = std::min(max_p_tmem, a->force_tiles_per_cta);
    // do not leave SMs idle: shrink P until the grid covers the machine once
    while (p_hi > 1 && (long long)((prm.n_tiles + p_hi - 1) / p_hi) * n_blocks < 148) --p_hi;
    static const int chunks[3] = {128, 64, 32};
    bool found = false;
    for (int want_stages = 3; want_stages >= 2 && !found; --want_stages) {
      for (int p = p_hi; p >= 1 && !found; --p) {
        for (int ci = 0; ci < 3 && !found; ++ci) {
          const int cb = chunks[ci];
          if (a->force_chunk_bytes && a->force_chunk_bytes != cb) continue;
          if (row_bytes % cb) continue;
          const int stage_bytes = (p * prm.box_rows + prm.b_taps * n_cta) * cb;
          int st = kSmemBudget / stage_bytes;
          const int total_steps = prm.n_loads * (row_bytes / cb);
          st = std::min(st, std::min(8, total_steps));
          if (a->force_stages) st = std::min(st, a->force_stages);
          if (st >= std::min(want_stages, total_steps) && st >= 1) {
            chunk = cb, P = p, stages = st, found = true;
          }
        }
      }
    }
    if (!found) return VK_E_UNSUPPORTED;
  }
  prm.tiles_per_cta = P;
  prm.k_chunks = row_bytes / chunk;
  prm.stages = stages;
  prm.tmem_cols = next_pow2_cols(P * prm.acc_stride);
  if (prm.tmem_cols > 512) return VK_E_UNSUPPORTED;

  // ---- epilogue ----
  prm.epi = a->epi;
  prm.ldo = a->ldo;
  prm.alpha = a->alpha;
  prm.round_out2 = a->round_out2;
  prm.bias = a->bias;
  prm.resid = a->resid;
  prm.mask = a->mask;
  prm.out1 = a->out1;
  prm.out2 = a->out2;
  prm.act_expclamp = a->act_expclamp;
  prm.clamp_lo = a->clamp_lo;
  prm.clamp_hi = a->clamp_hi;
  prm.crop_h = a->crop_h > 0 ? a->crop_h : prm.oh;
  prm.crop_w = a->crop_w > 0 ? a->crop_w : prm.ow;
  prm.cta_timing = a->cta_timing;

  // ---- tensor maps ----
  CUtensorMap ta, tb;
  {
    const uint64_t dims[4] = {uint64_t(a->ldx), uint64_t(a->iw), uint64_t(a->ih), uint64_t(a->n)};
    const uint64_t strides[3] = {uint64_t(row_bytes), uint64_t(row_bytes) * a->iw,
                                 uint64_t(row_bytes) * a->iw * a->ih};
    const uint32_t s = prm.a_stride;
    const uint32_t box[4] = {uint32_t(chunk / esize), uint32_t(tw) * s, uint32_t(box_h) * s, 1u};
    const uint32_t es[4] = {1u, s, s, 1u};
    int r = make_tensor_map(&ta, a->dtype, 4, a->x, dims, strides, box, es, chunk);
    if (r) return r;
  }
  {
    const uint64_t dims[3] = {uint64_t(a->ldx), uint64_t(a->wrows), uint64_t(taps)};
    const uint64_t strides[2] = {uint64_t(row_bytes), uint64_t(row_bytes) * a->wrows};
    const uint32_t box[3] = {uint32_t(chunk / esize), uint32_t(n_cta), 1u};
    const uint32_t es[3] = {1u, 1u, 1u};
    int r = make_tensor_map(&tb, a->dtype, 3, a->w, dims, strides, box, es, chunk);
    if (r) return r;
  }

  const int stage_bytes = (P * prm.box_rows + prm.b_taps * n_cta) * chunk;
  const int smem_bytes = stages * stage_bytes + 1024;
  dim3 grid((prm.n_tiles + P - 1) / P, n_blocks);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

#define VK_DISPATCH(DT)                                                        \
  switch (chunk) {                                                             \
    case 128: return launch_conv<DT, 128>(ta, tb, prm, grid, smem_bytes, st);  \
    case 64: return launch_conv<DT, 64>(ta, tb, prm, grid, smem_bytes, st);    \
    default: return launch_conv<DT, 32>(ta, tb, prm, grid, smem_bytes, st);    \
  }
  if (a->dtype == VK_BF16) {
    VK_DISPATCH(__nv_bfloat16)
  } else {
    VK_DISPATCH(float)
  }
#undef VK_DISPATCH
}

extern "C" const char* vk_version(void) { return "virnet_b200 0.1 (sm_100a; tcgen05+TMA)"; }
extern "C" uint64_t vk_launch_count(void) { return vk::g_launch_count.load(); }
extern "C" uint32_t vk_sizeof_conv_args(void) { return uint32_t(sizeof(vk_conv_args)); }
