// conv_v2_kernel instantiations: bf16_single (see vk_conv_v2_inst.inc)
#define VK_INST_DT __nv_bfloat16
#define VK_INST_PAIR false
#define VK_INST_NAME v2_launch_bf16_single
#include "vk_conv_v2_inst.inc"
