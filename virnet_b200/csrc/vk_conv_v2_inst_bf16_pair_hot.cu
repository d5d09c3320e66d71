// conv_v2_kernel instantiations specialised on the epilogue tensor combination (kEpi): bf16, CTA pairs — every
// (kernel configuration, combination) a denoising training step launches at its large layers (tools/v2_config_census.py).
#include <mutex>

#include "../../include/virnet_b200.h"
#include "vk_conv_v2_launch.h"
#include "vk_host.h"

namespace vk {
namespace {

template <int kChunk, int kNT, bool kFullK, int kEpi>
int launch_hot(const CUtensorMap& ta, const CUtensorMap& tb, const ConvV2Maps& em, const ConvV2Params& prm, int grid,
               int smem_bytes, cudaStream_t st) {
  static int cur = 0;
  static std::mutex mu;
  auto kern = conv_v2_kernel<__nv_bfloat16, kChunk, kNT, true, kFullK, kEpi>;
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

// returns VK_E_UNSUPPORTED when no specialised kernel exists for (chunk, nt, mode): the caller then uses the generic one
int v2_launch_bf16_pair_hot(int chunk, int nt, int mode, const CUtensorMap& ta, const CUtensorMap& tb, const ConvV2Maps& em,
                            const ConvV2Params& prm, int grid, int smem_bytes, cudaStream_t st) {
#define VK_HOT(C, T, M) \
  if (chunk == C && nt == T && mode == M && !prm.full_k) return launch_hot<C, T, false, M>(ta, tb, em, prm, grid, smem_bytes, st);
#define VK_HOT_MODES(C, T) VK_HOT(C, T, 5) VK_HOT(C, T, 6) VK_HOT(C, T, 7) VK_HOT(C, T, 8) VK_HOT(C, T, 14)
  VK_HOT_MODES(64, 9) VK_HOT_MODES(128, 9) VK_HOT_MODES(64, 3) VK_HOT_MODES(128, 3)
  // 16-channel-input layers (network heads, padded 3/4-channel images): SNet conv1, RNet head, tail dgrad
  VK_HOT(32, 9, 4) VK_HOT(32, 9, 5) VK_HOT(32, 9, 8) VK_HOT(32, 9, 12)
  // SFT-modulated residual blocks of the super-resolution network (mode + 16)
  VK_HOT(64, 9, 24) VK_HOT(64, 9, 28) VK_HOT(64, 9, 30) VK_HOT(64, 3, 24) VK_HOT(64, 3, 28) VK_HOT(64, 3, 30)
#undef VK_HOT_MODES
#undef VK_HOT
  // strided / transposed kinds (full-K mode): stride-2 conv (out1 + out2), ConvT (resid + out1 + out2), their dgrads
#define VK_HOT_FK(C, M) \
  if (chunk == C && mode == M && prm.full_k) return launch_hot<C, 1, true, M>(ta, tb, em, prm, grid, smem_bytes, st);
  VK_HOT_FK(64, 4) VK_HOT_FK(64, 6) VK_HOT_FK(64, 12) VK_HOT_FK(64, 14)
  VK_HOT_FK(128, 4) VK_HOT_FK(128, 6) VK_HOT_FK(128, 12) VK_HOT_FK(128, 14)
#undef VK_HOT_FK
  return VK_E_UNSUPPORTED;
}

}  // namespace vk
