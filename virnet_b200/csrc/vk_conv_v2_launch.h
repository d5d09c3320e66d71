// vk_conv_v2_launch.h — launch entry points of the persistent conv kernel, one per (dtype, pair mode);
// each is defined in its own translation unit (vk_conv_v2_inst_*.cu) so the 48 kernel instantiations build in parallel.
#pragma once
#include "vk_conv_v2.cuh"

namespace vk {
#define VK_V2_DECL(NAME)                                                                                              \
  int NAME(int chunk, int nt, const CUtensorMap& ta, const CUtensorMap& tb, const ConvV2Maps& em, const ConvV2Params& prm, \
           int grid, int smem_bytes, cudaStream_t st);
VK_V2_DECL(v2_launch_bf16_single)
VK_V2_DECL(v2_launch_bf16_pair)
VK_V2_DECL(v2_launch_tf32_single)
VK_V2_DECL(v2_launch_tf32_pair)
#undef VK_V2_DECL
// (chunk bytes, taps per weight item, epilogue mode) combinations of the specialised bf16 CTA-pair SLAB kernels — the
// single source of truth for vk_conv_v2_inst_bf16_pair_hot.cu (instantiation) and the host planner (staging layout)
constexpr bool v2_hot_slab_exists(int chunk, int nt, int mode) {
  const bool main_modes = mode == 5 || mode == 6 || mode == 7 || mode == 8 || mode == 14;
  const bool sft_modes = mode == 24 || mode == 28 || mode == 30;
  if ((chunk == 64 || chunk == 128) && (nt == 9 || nt == 3) && main_modes) return true;
  if (chunk == 32 && nt == 9 && (mode == 4 || mode == 5 || mode == 8 || mode == 12)) return true;
  if (chunk == 64 && (nt == 9 || nt == 3) && sft_modes) return true;
  return false;
}
// the same for the tf32 (fp32 storage) CTA-pair slab kernels of vk_conv_v2_inst_tf32_pair_hot.cu: what a tf32 training
// step launches (tools/v2_config_census.py 32 tf32)
constexpr bool v2_hot_slab_exists_tf32(int chunk, int nt, int mode) {
  const bool main_modes = mode == 5 || mode == 6 || mode == 7 || mode == 8 || mode == 14;
  if (chunk == 128 && (nt == 9 || nt == 3) && main_modes) return true;
  if (chunk == 32 && nt == 9 && (mode == 4 || mode == 5 || mode == 8 || mode == 12)) return true;
  return false;
}
// bf16 CTA-pair slab kernels specialised on the epilogue tensor combination `mode` (vk_conv_v2_inst_bf16_pair_hot.cu)
int v2_launch_bf16_pair_hot(int chunk, int nt, int mode, const CUtensorMap& ta, const CUtensorMap& tb, const ConvV2Maps& em,
                            const ConvV2Params& prm, int grid, int smem_bytes, cudaStream_t st);
int v2_launch_tf32_pair_hot(int chunk, int nt, int mode, const CUtensorMap& ta, const CUtensorMap& tb, const ConvV2Maps& em,
                            const ConvV2Params& prm, int grid, int smem_bytes, cudaStream_t st);
}  // namespace vk
