// conv_v2_kernel instantiations specialised on the epilogue tensor combination (kEpi): tf32 (fp32 storage), CTA pairs —
// every (kernel configuration, combination) the large layers of a denoising training step launch in the 1e-3-parity
// precision (tools/v2_config_census.py 32 tf32).  Round 1 ran this mode on the generic kernel, whose per-item epilogue
// is instruction-fetch bound.
#include <mutex>

#include "../../include/virnet_b200.h"
#include "vk_conv_v2_launch.h"
#include "vk_host.h"

namespace vk {
namespace {

template <int kChunk, int kNT, bool kFullK, int kEpi>
int launch_hot32(const CUtensorMap& ta, const CUtensorMap& tb, const ConvV2Maps& em, const ConvV2Params& prm, int grid,
                 int smem_bytes, cudaStream_t st) {
  static int cur = 0;
  static std::mutex mu;
  auto kern = conv_v2_kernel<float, kChunk, kNT, true, kFullK, kEpi>;
  {
    std::lock_guard<std::mutex> g(mu);
    if (smem_bytes > cur) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
      if (e != cudaSuccess) return int(e);
      cur = smem_bytes;
    }
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid), cfg.blockDim = dim3(v2_threads(true)), cfg.dynamicSmemBytes = smem_bytes, cfg.stream = st;
  cudaLaunchAttribute attr[2];
  cfg.attrs = attr, cfg.numAttrs = fill_launch_attrs(attr, true, prm.n_jobs <= 4LL * grid);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, em, prm);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return e != cudaSuccess ? int(e) : int(cudaGetLastError());
}

}  // namespace

int v2_launch_tf32_pair_hot(int chunk, int nt, int mode, const CUtensorMap& ta, const CUtensorMap& tb, const ConvV2Maps& em,
                            const ConvV2Params& prm, int grid, int smem_bytes, cudaStream_t st) {
#define VK_HOT(C, T, M) \
  if (chunk == C && nt == T && mode == M && !prm.full_k) return launch_hot32<C, T, false, M>(ta, tb, em, prm, grid, smem_bytes, st);
#define VK_HOT_MODES(C, T) VK_HOT(C, T, 5) VK_HOT(C, T, 6) VK_HOT(C, T, 7) VK_HOT(C, T, 8) VK_HOT(C, T, 14)
  VK_HOT_MODES(128, 9) VK_HOT_MODES(128, 3)
  VK_HOT(32, 9, 4) VK_HOT(32, 9, 5) VK_HOT(32, 9, 8) VK_HOT(32, 9, 12)
#undef VK_HOT_MODES
#undef VK_HOT
#define VK_HOT_FK(C, M) \
  if (chunk == C && mode == M && prm.full_k) return launch_hot32<C, 1, true, M>(ta, tb, em, prm, grid, smem_bytes, st);
  VK_HOT_FK(128, 4) VK_HOT_FK(128, 6) VK_HOT_FK(128, 12) VK_HOT_FK(128, 14)
#undef VK_HOT_FK
  return VK_E_UNSUPPORTED;
}

}  // namespace vk
