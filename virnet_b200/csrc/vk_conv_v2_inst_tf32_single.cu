// conv_v2_kernel instantiations: tf32_single (see vk_conv_v2_inst.inc)
#define VK_INST_DT float
#define VK_INST_PAIR false
#define VK_INST_NAME v2_launch_tf32_single
#include "vk_conv_v2_inst.inc"
