// vk_host.h — host-side declarations shared by the .cu translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdlib>

namespace vk {

extern std::atomic<uint64_t> g_launch_count;

// Encode (or fetch from the cache) a tiled TMA descriptor.  dtype is VK_BF16 /
// VK_TF32 (fp32 storage).  dims[0] is the contiguous dimension; strides_bytes
// has rank-1 entries (dims 1..rank-1).  swizzle_bytes: 0/32/64/128, or 129 for
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  Returns 0 or a VK_E_* code.
int make_tensor_map(CUtensorMap* out, int dtype, int rank, const void* ptr, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estrides,
                    int swizzle_bytes);

// Launch attributes of the persistent tcgen05 kernels: CTA-pair clusters and programmatic dependent launch (PDL).
// With PDL the CTAs of kernel N+1 become resident as the CTAs of kernel N retire and run their prologue (barrier init,
// TMEM allocation, descriptor prefetch) under N's tail; every such kernel executes griddepcontrol.wait before it touches
// global memory, which blocks until N has completed and flushed.  VK_NO_PDL=1 restores plain stream order.
inline bool pdl_enabled() {
  static const bool on = std::getenv("VK_NO_PDL") == nullptr;
  return on;
}
// `small`: the launch is short (a few jobs per CTA).  Only then is PDL requested: a dependent kernel becomes resident on
// the SMs its predecessor has left and sits there until the predecessor completes, which at large batches locks out the
// weight-gradient kernels of the side stream that would otherwise fill exactly those SMs (measured at batch 32, same
// box: 9.49 ms without PDL, 9.62 ms with; at 2 patches per GPU PDL is worth 17 %).
inline unsigned fill_launch_attrs(cudaLaunchAttribute* attr, bool cluster2, bool small) {
  unsigned n = 0;
  if (cluster2) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = 2, attr[n].val.clusterDim.y = 1, attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (small && pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  return n;
}

}  // namespace vk
