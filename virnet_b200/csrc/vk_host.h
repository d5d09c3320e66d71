// vk_host.h — host-side declarations shared by the .cu translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace vk {

extern std::atomic<uint64_t> g_launch_count;

// Encode (or fetch from the cache) a tiled TMA descriptor.  dtype is VK_BF16 /
// VK_TF32 (fp32 storage).  dims[0] is the contiguous dimension; strides_bytes
// has rank-1 entries (dims 1..rank-1).  swizzle_bytes: 0/32/64/128, or 129 for
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  Returns 0 or a VK_E_* code.
int make_tensor_map(CUtensorMap* out, int dtype, int rank, const void* ptr, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estrides,
                    int swizzle_bytes);

}  // namespace vk
