// conv_v2_kernel instantiations specialised on the epilogue tensor combination (kEpi = 5, 7, 8, 14): bf16, CTA pairs,
// slab mode — the residual-block convolutions of a training step (conv1 / conv2, forward and dgrad).
#include <mutex>

#include "../../include/virnet_b200.h"
#include "vk_conv_v2_launch.h"
#include "vk_host.h"

namespace vk {
namespace {

template <int kChunk, int kNT, int kEpi>
int launch_hot(const CUtensorMap& ta, const CUtensorMap& tb, const ConvV2Maps& em, const ConvV2Params& prm, int grid,
               int smem_bytes, cudaStream_t st) {
  static int cur = 0;
  static std::mutex mu;
  auto kern = conv_v2_kernel<__nv_bfloat16, kChunk, kNT, true, false, kEpi>;
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
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, em, prm);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return e != cudaSuccess ? int(e) : int(cudaGetLastError());
}

}  // namespace

// returns VK_E_UNSUPPORTED when no specialised kernel exists for (chunk, nt, mode): the caller then uses the generic one
int v2_launch_bf16_pair_hot(int chunk, int nt, int mode, const CUtensorMap& ta, const CUtensorMap& tb, const ConvV2Maps& em,
                            const ConvV2Params& prm, int grid, int smem_bytes, cudaStream_t st) {
#define VK_HOT(C, T, M) \
  if (chunk == C && nt == T && mode == M) return launch_hot<C, T, M>(ta, tb, em, prm, grid, smem_bytes, st);
#define VK_HOT_MODES(C, T) VK_HOT(C, T, 5) VK_HOT(C, T, 7) VK_HOT(C, T, 8) VK_HOT(C, T, 14)
  VK_HOT_MODES(64, 9) VK_HOT_MODES(128, 9) VK_HOT_MODES(64, 3) VK_HOT_MODES(128, 3)
#undef VK_HOT_MODES
#undef VK_HOT
  return VK_E_UNSUPPORTED;
}

}  // namespace vk
