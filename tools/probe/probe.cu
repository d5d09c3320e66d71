// probe.cu — stand-alone B200 micro-probes used to make design decisions for the conv kernel
// (not part of the product library).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe probe.cu
//
//   ./probe tma    : per-SM and chip-wide TMA load throughput vs box inner width / swizzle / box shape
//   ./probe shift  : is a UMMA A-operand descriptor whose start address sits at an arbitrary ROW offset
//                    inside a swizzled TMA slab read correctly (with / without the base_offset field)?
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>

#include "../../virnet_b200/csrc/vk_common.cuh"

using namespace vk;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn enc() {
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    fn = reinterpret_cast<EncodeFn>(p);
  }
  return fn;
}
static CUtensorMap make_map(void* ptr, int rank, const uint64_t* dims, const uint64_t* strides, const uint32_t* box,
                            int swizzle) {
  CUtensorMap m;
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) gd[i] = dims[i], bx[i] = box[i], es[i] = 1;
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides[i];
  CUtensorMapSwizzle sw = swizzle == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                          : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, ptr, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("encode failed %d\n", int(r));
    exit(1);
  }
  return m;
}

// =====================================================================================
// TMA throughput
// =====================================================================================
struct TmaCfg {
  int inner_elems, bw, bh;   // box
  int nchunk;                // channel chunks (c coordinate cycles over nchunk * inner_elems)
  int W, H, N;               // tensor
  int stages, iters;
  int coff;                  // channel offset of chunk 0
};

__global__ void __launch_bounds__(128, 1) tma_bw_kernel(const __grid_constant__ CUtensorMap tm, TmaCfg c,
                                                        long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[16];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int box_bytes = c.inner_elems * 2 * c.bw * c.bh;
  const int stage_stride = (box_bytes + 1023) / 1024 * 1024;
  if (threadIdx.x == 0) {
    for (int s = 0; s < c.stages; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int tx_n = c.W / c.bw, ty_n = c.H / c.bh;
    const int tiles = tx_n * ty_n * c.N;
    long long t0 = clock64();
    for (int it = 0; it < c.iters + c.stages; ++it) {
      const int s = it % c.stages;
      if (it >= c.stages) mbar_wait(&full[s], ((it / c.stages) - 1) & 1);
      if (it < c.iters) {
        const int t = (blockIdx.x * 131 + it / c.nchunk) % tiles;
        const int ch = it % c.nchunk;
        const int img = t / (tx_n * ty_n), r = t % (tx_n * ty_n);
        mbar_arrive_expect_tx(&full[s], box_bytes);
        tma_load_4d(smem + s * stage_stride, &tm, &full[s], c.coff + ch * c.inner_elems, (r % tx_n) * c.bw - 1,
                    (r / tx_n) * c.bh - 1, img);
      }
    }
    long long t1 = clock64();
    cycles_out[blockIdx.x] = t1 - t0;
  }
}

static void run_tma(const char* name, int C, int inner_elems, int swz, int bw, int bh, int nchunk, int stages,
                    int nimg, int coff = 0) {
  TmaCfg c{};
  c.inner_elems = inner_elems, c.bw = bw, c.bh = bh, c.nchunk = nchunk;
  c.W = 128, c.H = 128, c.N = nimg, c.stages = stages, c.iters = 3000, c.coff = coff;
  size_t bytes = size_t(nimg) * 128 * 128 * C * 2;
  void* d;
  CK(cudaMalloc(&d, bytes));
  CK(cudaMemset(d, 0, bytes));
  uint64_t dims[4] = {uint64_t(C), 128, 128, uint64_t(nimg)};
  uint64_t strides[3] = {uint64_t(C) * 2, uint64_t(C) * 2 * 128, uint64_t(C) * 2 * 128 * 128};
  uint32_t box[4] = {uint32_t(inner_elems), uint32_t(bw), uint32_t(bh), 1};
  CUtensorMap tm = make_map(d, 4, dims, strides, box, swz);
  long long* cyc;
  CK(cudaMalloc(&cyc, 148 * 8));
  const int box_bytes = inner_elems * 2 * bw * bh;
  const int smem = stages * ((box_bytes + 1023) / 1024 * 1024) + 1024;
  CK(cudaFuncSetAttribute(tma_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int grid : {1, 148}) {
    float best = 1e9;
    long long cy = 0;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0), cudaEventCreate(&e1);
      cudaEventRecord(e0);
      tma_bw_kernel<<<grid, 128, std::max(smem, 190 * 1024)>>>(tm, c, cyc);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) {
        best = ms;
        std::vector<long long> h(grid);
        CK(cudaMemcpy(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost));
        cy = 0;
        for (auto v : h) cy = std::max(cy, v);
      }
    }
    double total = double(c.iters) * box_bytes * grid;
    printf("%-44s grid=%3d box=%6dB stages=%d in-flight=%3dKB : %6.1f B/clk/SM  %7.1f GB/s total (%.3f ms, %lld cyc)\n",
           name, grid, box_bytes, stages, stages * box_bytes / 1024, double(c.iters) * box_bytes / cy,
           total / (best * 1e-3) / 1e9, best, cy);
  }
  cudaFree(d);
  cudaFree(cyc);
}


// batch mode: arm ONE barrier with the bytes of `batch` boxes, issue them back to back, wait once.
struct Tma2Cfg {
  int rank, inner_elems, rows, bw, bh, batch, iters, tiles, issuers;
};
__global__ void __launch_bounds__(128, 1) tma_batch_kernel(const __grid_constant__ CUtensorMap tm, Tma2Cfg c,
                                                           long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[4];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int box_bytes = c.inner_elems * 2 * c.rows;
  const int stride = (box_bytes + 1023) / 1024 * 1024;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&full[i], 1);
    fence_barrier_init();
    tma_prefetch_desc(&tm);
  }
  __syncthreads();
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && w < c.issuers) {
    long long t0 = clock64();
    for (int it = 0; it < c.iters; ++it) {
      mbar_arrive_expect_tx(&full[w], box_bytes * c.batch);
      for (int b = 0; b < c.batch; ++b) {
        const int t = (blockIdx.x * 131 + (it * c.batch + b) * 7 + w * 3) % c.tiles;
        uint8_t* dst = smem + (w * c.batch + b) * stride;
        if (c.rank == 2) {
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
              ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(&full[w])), "r"(0), "r"(t * c.rows)
              : "memory");
        } else {
          const int tx_n = 128 / c.bw, ty_n = 128 / c.bh;
          const int img = t / (tx_n * ty_n), r = t % (tx_n * ty_n);
          if (c.rank == 3) tma_load_3d(dst, &tm, &full[w], 0, (r % tx_n) * c.bw, img * 128 + (r / tx_n) * c.bh);
          else tma_load_4d(dst, &tm, &full[w], 0, (r % tx_n) * c.bw, (r / tx_n) * c.bh, img);
        }
      }
      mbar_wait(&full[w], it & 1);
    }
    long long t1 = clock64();
    if (w == 0) cycles_out[blockIdx.x] = t1 - t0;
  }
}

static void run_tma_batch(const char* name, int rank, int C, int inner_elems, int swz, int bw, int bh, int batch,
                          int issuers, int nimg) {
  Tma2Cfg c{};
  c.rank = rank, c.inner_elems = inner_elems, c.rows = bw * bh, c.bw = bw, c.bh = bh, c.batch = batch, c.iters = 400;
  c.issuers = issuers;
  size_t bytes = size_t(nimg) * 128 * 128 * C * 2;
  void* d;
  CK(cudaMalloc(&d, bytes));
  CK(cudaMemset(d, 0, bytes));
  CUtensorMap tm;
  if (rank == 2) {
    uint64_t dims[2] = {uint64_t(C), uint64_t(nimg) * 128 * 128};
    uint64_t strides[1] = {uint64_t(C) * 2};
    uint32_t box[2] = {uint32_t(inner_elems), uint32_t(c.rows)};
    tm = make_map(d, 2, dims, strides, box, swz);
    c.tiles = nimg * 128 * 128 / c.rows;
  } else if (rank == 3) {
    uint64_t dims[3] = {uint64_t(C), 128, uint64_t(nimg) * 128};
    uint64_t strides[2] = {uint64_t(C) * 2, uint64_t(C) * 2 * 128};
    uint32_t box[3] = {uint32_t(inner_elems), uint32_t(bw), uint32_t(bh)};
    tm = make_map(d, 3, dims, strides, box, swz);
    c.tiles = nimg * (128 / bw) * (128 / bh);
  } else {
    uint64_t dims[4] = {uint64_t(C), 128, 128, uint64_t(nimg)};
    uint64_t strides[3] = {uint64_t(C) * 2, uint64_t(C) * 2 * 128, uint64_t(C) * 2 * 128 * 128};
    uint32_t box[4] = {uint32_t(inner_elems), uint32_t(bw), uint32_t(bh), 1};
    tm = make_map(d, 4, dims, strides, box, swz);
    c.tiles = nimg * (128 / bw) * (128 / bh);
  }
  long long* cyc;
  CK(cudaMalloc(&cyc, 148 * 8));
  CK(cudaFuncSetAttribute(tma_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int box_bytes = inner_elems * 2 * c.rows;
  for (int grid : {1, 148}) {
    long long cy = 1LL << 60;
    for (int rep = 0; rep < 3; ++rep) {
      tma_batch_kernel<<<grid, 128, 200 * 1024>>>(tm, c, cyc);
      CK(cudaDeviceSynchronize());
      std::vector<long long> h(grid);
      CK(cudaMemcpy(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost));
      long long m = 0;
      for (auto v : h) m = std::max(m, v);
      cy = std::min(cy, m);
    }
    printf("%-40s rank=%d grid=%3d box=%6dB(%3d rows) batch=%2d issuers=%d : %7.0f cyc/batch %6.0f cyc/box %6.1f B/clk/SM\n", name,
           rank, grid, box_bytes, c.rows, batch, issuers, double(cy) / c.iters, double(cy) / c.iters / batch,
           double(c.iters) * batch * issuers * box_bytes / cy);
  }
  cudaFree(d);
  cudaFree(cyc);
}


// steady-state ring: `stages` boxes in flight, rank 2 / 3 / 4 views of the same NHWC tensor
struct Tma3Cfg {
  int rank, inner_elems, bw, bh, stages, iters, halo, W, H, N;
};
__global__ void __launch_bounds__(128, 1) tma_ring_kernel(const __grid_constant__ CUtensorMap tm, Tma3Cfg c,
                                                          long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[16];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int box_bytes = c.inner_elems * 2 * c.bw * c.bh;
  const int stride = (box_bytes + 1023) / 1024 * 1024;
  if (threadIdx.x == 0) {
    for (int s = 0; s < c.stages; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
    tma_prefetch_desc(&tm);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int tx_n = c.W / c.bw, ty_n = c.H / c.bh;
    const int tiles = tx_n * ty_n * c.N;
    long long t0 = clock64();
    for (int it = 0; it < c.iters + c.stages; ++it) {
      const int s = it % c.stages;
      if (it >= c.stages) mbar_wait(&full[s], ((it / c.stages) - 1) & 1);
      if (it < c.iters) {
        const int t = (blockIdx.x * 131 + it * 7) % tiles;
        const int img = t / (tx_n * ty_n), r = t % (tx_n * ty_n);
        const int x0 = (r % tx_n) * c.bw - c.halo, y0 = (r / tx_n) * c.bh - c.halo;
        uint8_t* dst = smem + s * stride;
        mbar_arrive_expect_tx(&full[s], box_bytes);
        if (c.rank == 2) {
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
              ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(&full[s])), "r"(0),
              "r"((img * c.H + (r / tx_n) * c.bh) * c.W + (r % tx_n) * c.bw)
              : "memory");
        } else if (c.rank == 3) {
          tma_load_3d(dst, &tm, &full[s], 0, x0, img * c.H + (r / tx_n) * c.bh - c.halo);
        } else {
          tma_load_4d(dst, &tm, &full[s], 0, x0, y0, img);
        }
      }
    }
    long long t1 = clock64();
    cycles_out[blockIdx.x] = t1 - t0;
  }
}
static void run_tma_ring(const char* name, int rank, int C, int inner_elems, int swz, int bw, int bh, int stages,
                         int halo, int nimg, int promo = 256) {
  Tma3Cfg c{};
  c.rank = rank, c.inner_elems = inner_elems, c.bw = bw, c.bh = bh, c.stages = stages, c.iters = 2000, c.halo = halo;
  c.W = 128, c.H = 128, c.N = nimg;
  size_t bytes = size_t(nimg) * 128 * 128 * C * 2;
  void* d;
  CK(cudaMalloc(&d, bytes));
  CK(cudaMemset(d, 0, bytes));
  CUtensorMap tm;
  const uint64_t rb = uint64_t(C) * 2;
  if (rank == 2) {
    uint64_t dims[2] = {uint64_t(C), uint64_t(nimg) * 128 * 128};
    uint64_t strides[1] = {rb};
    uint32_t box[2] = {uint32_t(inner_elems), uint32_t(bw * bh)};
    tm = make_map(d, 2, dims, strides, box, swz);
  } else if (rank == 3) {
    uint64_t dims[3] = {uint64_t(C), 128, uint64_t(nimg) * 128};
    uint64_t strides[2] = {rb, rb * 128};
    uint32_t box[3] = {uint32_t(inner_elems), uint32_t(bw), uint32_t(bh)};
    tm = make_map(d, 3, dims, strides, box, swz);
  } else {
    uint64_t dims[4] = {uint64_t(C), 128, 128, uint64_t(nimg)};
    uint64_t strides[3] = {rb, rb * 128, rb * 128 * 128};
    uint32_t box[4] = {uint32_t(inner_elems), uint32_t(bw), uint32_t(bh), 1};
    tm = make_map(d, 4, dims, strides, box, swz);
  }
  long long* cyc;
  CK(cudaMalloc(&cyc, 148 * 8));
  CK(cudaFuncSetAttribute(tma_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
  const int box_bytes = inner_elems * 2 * bw * bh;
  for (int grid : {1, 148}) {
    long long cy = 1LL << 60;
    for (int rep = 0; rep < 3; ++rep) {
      tma_ring_kernel<<<grid, 128, 210 * 1024>>>(tm, c, cyc);
      CK(cudaDeviceSynchronize());
      std::vector<long long> h(grid);
      CK(cudaMemcpy(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost));
      long long m = 0;
      for (auto v : h) m = std::max(m, v);
      cy = std::min(cy, m);
    }
    printf("%-34s rank=%d grid=%3d box=%2dx%-2d %6dB stages=%2d halo=%d : %6.0f cyc/box %5.2f cyc/row %6.1f B/clk/SM\n", name, rank,
           grid, bw, bh, box_bytes, stages, halo, double(cy) / c.iters, double(cy) / c.iters / (bw * bh),
           double(c.iters) * box_bytes / cy);
  }
  cudaFree(d);
  cudaFree(cyc);
}


// mode 0: `issuers` WARPS (lane 0 of each) issue 4D boxes; mode 1: `issuers` LANES of warp 0 issue 4D boxes in one
// instruction; mode 2: 1D bulk copies (cp.async.bulk.shared::cluster.global) of `bytes1d` bytes from lane 0 of `issuers` warps.
struct Tma5Cfg {
  int mode, issuers, batch, iters, bw, bh, inner_elems, tiles, bytes1d;
};
__global__ void __launch_bounds__(256, 1) tma_multi_kernel(const __grid_constant__ CUtensorMap tm, const uint8_t* src,
                                                           Tma5Cfg c, long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[8];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int box_bytes = c.mode == 2 ? c.bytes1d : c.inner_elems * 2 * c.bw * c.bh;
  const int stride = (box_bytes + 1023) / 1024 * 1024;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&full[i], c.mode == 1 ? 1 : 1);
    fence_barrier_init();
    tma_prefetch_desc(&tm);
  }
  __syncthreads();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tx_n = 128 / c.bw, ty_n = 128 / c.bh;
  if (c.mode == 1) {
    if (w == 0) {
      long long t0 = clock64();
      for (int it = 0; it < c.iters; ++it) {
        if (lane == 0) mbar_arrive_expect_tx(&full[0], box_bytes * c.batch * c.issuers);
        __syncwarp();
        for (int b = 0; b < c.batch; ++b) {
          if (lane < c.issuers) {
            const int t = (blockIdx.x * 131 + ((it * c.batch + b) * c.issuers + lane) * 7) % c.tiles;
            const int img = t / (tx_n * ty_n), r = t % (tx_n * ty_n);
            tma_load_4d(smem + (b * c.issuers + lane) * stride, &tm, &full[0], 0, (r % tx_n) * c.bw, (r / tx_n) * c.bh, img);
          }
        }
        mbar_wait(&full[0], it & 1);
        __syncwarp();
      }
      long long t1 = clock64();
      if (lane == 0) cycles_out[blockIdx.x] = t1 - t0;
    }
    return;
  }
  if (lane == 0 && w < c.issuers) {
    long long t0 = clock64();
    for (int it = 0; it < c.iters; ++it) {
      mbar_arrive_expect_tx(&full[w], box_bytes * c.batch);
      for (int b = 0; b < c.batch; ++b) {
        uint8_t* dst = smem + (w * c.batch + b) * stride;
        if (c.mode == 0) {
          const int t = (blockIdx.x * 131 + ((it * c.batch + b) * c.issuers + w) * 7) % c.tiles;
          const int img = t / (tx_n * ty_n), r = t % (tx_n * ty_n);
          tma_load_4d(dst, &tm, &full[w], 0, (r % tx_n) * c.bw, (r / tx_n) * c.bh, img);
        } else {
          const size_t off = (size_t((blockIdx.x * 131 + ((it * c.batch + b) * c.issuers + w) * 7) % 1024)) * 32768;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           smem_u32(dst)),
                       "l"(src + off), "r"(box_bytes), "r"(smem_u32(&full[w]))
                       : "memory");
        }
      }
      mbar_wait(&full[w], it & 1);
    }
    long long t1 = clock64();
    if (w == 0) cycles_out[blockIdx.x] = t1 - t0;
  }
}
static void run_tma_multi(const char* name, int mode, int issuers, int batch, int bw, int bh, int bytes1d) {
  Tma5Cfg c{};
  const int C = 192, nimg = 8;
  c.mode = mode, c.issuers = issuers, c.batch = batch, c.iters = 300, c.bw = bw, c.bh = bh, c.inner_elems = 64;
  c.tiles = nimg * (128 / bw) * (128 / bh), c.bytes1d = bytes1d;
  size_t bytes = size_t(nimg) * 128 * 128 * C * 2;
  void* d;
  CK(cudaMalloc(&d, bytes));
  CK(cudaMemset(d, 0, bytes));
  uint64_t dims[4] = {uint64_t(C), 128, 128, uint64_t(nimg)};
  uint64_t strides[3] = {uint64_t(C) * 2, uint64_t(C) * 2 * 128, uint64_t(C) * 2 * 128 * 128};
  uint32_t box[4] = {64, uint32_t(bw), uint32_t(bh), 1};
  CUtensorMap tm = make_map(d, 4, dims, strides, box, 128);
  long long* cyc;
  CK(cudaMalloc(&cyc, 148 * 8));
  CK(cudaFuncSetAttribute(tma_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
  const int box_bytes = mode == 2 ? bytes1d : 64 * 2 * bw * bh;
  for (int grid : {1, 148}) {
    long long cy = 1LL << 60;
    for (int rep = 0; rep < 3; ++rep) {
      tma_multi_kernel<<<grid, 256, 210 * 1024>>>(tm, reinterpret_cast<const uint8_t*>(d), c, cyc);
      CK(cudaDeviceSynchronize());
      std::vector<long long> h(grid);
      CK(cudaMemcpy(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost));
      long long m = 0;
      for (auto v : h) m = std::max(m, v);
      cy = std::min(cy, m);
    }
    const double nbox = double(c.iters) * batch * issuers;
    printf("%-40s mode=%d grid=%3d box=%6dB issuers=%d batch=%d : %6.0f cyc per box (all issuers) %6.1f B/clk/SM\n", name, mode, grid,
           box_bytes, issuers, batch, cy / nbox, nbox * box_bytes / cy);
  }
  cudaFree(d);
  cudaFree(cyc);
}

// =====================================================================================
// UMMA row-shift probe
// =====================================================================================
__device__ __forceinline__ uint64_t make_desc_bo(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout,
                                                 uint32_t base_off) {
  uint64_t d = make_smem_desc(saddr, lbo, sbo, layout);
  d |= static_cast<uint64_t>(base_off & 7u) << 49;
  return d;
}

struct ShiftCfg {
  int chunk_bytes;   // 64 or 128 (swizzle width == row pitch)
  int slab_w, slab_h;
  int row_shift;     // start row inside the slab
  int sbo_rows;      // rows between 8-row groups
  int bo_mode;       // 0: base_offset = 0; 1: (start_addr >> 7) & 7
  int n;             // GEMM N
};

__global__ void __launch_bounds__(128, 1) shift_kernel(const __grid_constant__ CUtensorMap tma,
                                                       const __grid_constant__ CUtensorMap tmb, ShiftCfg c,
                                                       float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full, done;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int a_bytes = c.slab_w * c.slab_h * c.chunk_bytes;
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + (a_bytes + 1023) / 1024 * 1024;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&full, 1);
    mbar_init(&done, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 64);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&full, a_bytes + c.n * c.chunk_bytes);
    tma_load_4d(a_s, &tma, &full, 0, 0, 0, 0);
    tma_load_3d(b_s, &tmb, &full, 0, 0, 0);
    mbar_wait(&full, 0);
    tc_fence_after_sync();
    const uint32_t layout = layout_type_for_swizzle(c.chunk_bytes);
    const uint32_t idesc = make_idesc(1, 128, c.n, 0, 0);
    const uint32_t a0 = smem_u32(a_s) + c.row_shift * c.chunk_bytes;
    const uint32_t b0 = smem_u32(b_s);
    for (int k = 0; k < c.chunk_bytes / 32; ++k) {
      const uint32_t aa = a0 + k * 32;
      const uint32_t bo = c.bo_mode ? ((aa >> 7) & 7u) : 0u;
      const uint64_t ad = make_desc_bo(aa, 16, c.sbo_rows * c.chunk_bytes, layout, bo);
      const uint64_t bd = make_desc_bo(b0 + k * 32, 16, 8 * c.chunk_bytes, layout, 0);
      umma_ss<false>(tmem, ad, bd, idesc, k > 0);
    }
    umma_commit(&done);
  }
  mbar_wait(&done, 0);
  tc_fence_after_sync();
  const int lane = threadIdx.x & 31;
  for (int jc = 0; jc < c.n; jc += 16) {
    uint32_t rr[16];
    __syncwarp();
    tmem_ld16(tmem + (uint32_t(warp * 32) << 16) + jc, rr);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * c.n + jc + i] = __uint_as_float(rr[i]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem, 64);
  }
}

static float bf16_round(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}
static uint16_t bf16_bits(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  return uint16_t(u >> 16);
}

static void run_shift() {
  const int Wimg = 32, Himg = 32, N = 64;
  for (int chunk : {64, 128}) {
    const int kel = chunk / 2;
    // activation tensor [Himg][Wimg][kel] and weights [N][kel]
    std::vector<float> xa(size_t(Himg) * Wimg * kel), wb(size_t(N) * kel);
    std::vector<uint16_t> xa16(xa.size()), wb16(wb.size());
    srand(1);
    for (size_t i = 0; i < xa.size(); ++i) xa[i] = bf16_round((rand() % 2001 - 1000) / 1000.f), xa16[i] = bf16_bits(xa[i]);
    for (size_t i = 0; i < wb.size(); ++i) wb[i] = bf16_round((rand() % 2001 - 1000) / 1000.f), wb16[i] = bf16_bits(wb[i]);
    void *dx, *dw;
    float* dout;
    CK(cudaMalloc(&dx, xa16.size() * 2));
    CK(cudaMalloc(&dw, wb16.size() * 2));
    CK(cudaMalloc(&dout, 128 * N * 4));
    CK(cudaMemcpy(dx, xa16.data(), xa16.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw, wb16.data(), wb16.size() * 2, cudaMemcpyHostToDevice));
    for (int geom = 0; geom < 2; ++geom) {
      // geom 0: tile 8 wide x 16 tall, slab 10 x 18 (SBO = 10 rows);  geom 1: tile 16 x 8, slab 16 x 10 (SBO = 8 rows, v1 layout)
      const int tw = geom == 0 ? 8 : 16, th = 128 / tw;
      const int sw_ = geom == 0 ? tw + 2 : tw, sh_ = th + 2;
      uint64_t dims[4] = {uint64_t(kel), uint64_t(Wimg), uint64_t(Himg), 1};
      uint64_t strides[3] = {uint64_t(kel) * 2, uint64_t(kel) * 2 * Wimg, uint64_t(kel) * 2 * Wimg * Himg};
      uint32_t box[4] = {uint32_t(kel), uint32_t(sw_), uint32_t(sh_), 1};
      CUtensorMap tma = make_map(dx, 4, dims, strides, box, chunk);
      uint64_t bd[3] = {uint64_t(kel), uint64_t(N), 1};
      uint64_t bs[2] = {uint64_t(kel) * 2, uint64_t(kel) * 2 * N};
      uint32_t bb[3] = {uint32_t(kel), uint32_t(N), 1};
      CUtensorMap tmb = make_map(dw, 3, bd, bs, bb, chunk);
      CK(cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      for (int r = 0; r < 3; ++r)
        for (int s = 0; s < (geom == 0 ? 3 : 1); ++s)
          for (int bo = 0; bo < 2; ++bo) {
            ShiftCfg c{};
            c.chunk_bytes = chunk, c.slab_w = sw_, c.slab_h = sh_, c.row_shift = r * sw_ + s;
            c.sbo_rows = geom == 0 ? sw_ : 8, c.bo_mode = bo, c.n = N;
            CK(cudaMemset(dout, 0, 128 * N * 4));
            shift_kernel<<<1, 128, 100 * 1024>>>(tma, tmb, c, dout);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
              printf("chunk=%d geom=%d r=%d s=%d bo=%d : CUDA error %s\n", chunk, geom, r, s, bo, cudaGetErrorString(e));
              exit(1);
            }
            std::vector<float> h(128 * N);
            CK(cudaMemcpy(h.data(), dout, h.size() * 4, cudaMemcpyDeviceToHost));
            double maxerr = 0;
            for (int i = 0; i < 128; ++i) {
              const int y = i / tw + r, x = i % tw + s;   // slab origin = image (0,0)
              for (int n = 0; n < N; ++n) {
                double acc = 0;
                for (int k = 0; k < kel; ++k) acc += double(xa[(size_t(y) * Wimg + x) * kel + k]) * wb[size_t(n) * kel + k];
                maxerr = std::max(maxerr, std::fabs(acc - h[i * N + n]));
              }
            }
            printf("shift probe: swizzle=%3d tile=%2dx%-2d slab_w=%2d tap(r=%d,s=%d) start_row=%3d base_offset=%s : max|err| = %.3e %s\n",
                   chunk, tw, th, sw_, r, s, c.row_shift, bo ? "(addr>>7)&7" : "0", maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
          }
    }
    cudaFree(dx), cudaFree(dw), cudaFree(dout);
  }
}


// =====================================================================================
// UMMA issue-rate probe: operands resident in smem (garbage data), one thread issues `iters` batches of
// `per_commit` MMAs (M=128, N=n, K=16 bf16) round-robin over `nacc` accumulators, commit + wait per batch.
// =====================================================================================
struct MmaCfg {
  int n, chunk_bytes, a_sbo_rows, per_commit, iters, nacc, a_rows_total, wait_each;
};
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(MmaCfg c, long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t done[2];
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&done[0], 1);
    mbar_init(&done[1], 1);
    fence_barrier_init();
  }
  // zero the operand region so no NaN slow paths are involved
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t layout = layout_type_for_swizzle(c.chunk_bytes);
    const uint32_t idesc = make_idesc(1, 128, c.n, 0, 0);
    const uint32_t a0 = smem_u32(smem);
    const uint32_t b0 = a0 + 64 * 1024;
    const int acc_stride = (c.n + 31) / 32 * 32;
    const uint64_t a_hi = make_smem_desc(0, 16, c.a_sbo_rows * c.chunk_bytes, layout) & 0xFFFFFFFF00000000ull;
    const uint64_t b_hi = make_smem_desc(0, 16, 8 * c.chunk_bytes, layout) & 0xFFFFFFFF00000000ull;
    const uint32_t a_lo = ((a0 & 0x3FFFFu) >> 4) | (1u << 16), b_lo = ((b0 & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t btap16 = uint32_t(c.n * c.chunk_bytes) >> 4;
    const uint32_t arow16 = uint32_t(c.chunk_bytes) >> 4;
    const uint32_t aslab16 = uint32_t(c.a_sbo_rows * c.chunk_bytes) >> 4;
    long long t0 = clock64();
    for (int it = 0; it < c.iters; ++it) {
      // per_commit = 36: 3 tap rows x 3 taps x 4 k-steps (2 k-steps for 64-byte chunks, issued twice)
      const uint32_t d = tmem + (it % c.nacc) * acc_stride;
      for (int m = 0; m < c.per_commit; m += 12) {
        const uint32_t tr = (m / 12) % 3;
        const uint32_t ar = a_lo + tr * aslab16, br = b_lo;   // B: 3 distinct blocks only (smem budget)
#pragma unroll
        for (int s3 = 0; s3 < 3; ++s3) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t kk = c.chunk_bytes == 128 ? k : (k & 1);
            umma_ss<false>(d, a_hi | (ar + s3 * arow16 + 2 * kk), b_hi | (br + s3 * btap16 + 2 * kk), idesc, 1);
          }
        }
      }
      umma_commit(&done[it & 1]);
      if (c.wait_each) mbar_wait(&done[it & 1], (it >> 1) & 1);
      else if (it >= 1) mbar_wait(&done[(it - 1) & 1], ((it - 1) >> 1) & 1);
    }
    if (!c.wait_each) mbar_wait(&done[(c.iters - 1) & 1], ((c.iters - 1) >> 1) & 1);
    long long t1 = clock64();
    cycles_out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem, 512);
  }
}
static void run_mma(const char* name, int n, int chunk, int sbo_rows, int per_commit, int nacc, int wait_each = 0) {
  MmaCfg c{};
  c.n = n, c.chunk_bytes = chunk, c.a_sbo_rows = sbo_rows, c.per_commit = per_commit, c.iters = 200, c.nacc = nacc;
  c.wait_each = wait_each;
  long long* cyc;
  CK(cudaMalloc(&cyc, 148 * 8));
  CK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int grid : {1, 148}) {
    long long cy = 1LL << 60;
    float best = 1e9;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0), cudaEventCreate(&e1);
      cudaEventRecord(e0);
      mma_rate_kernel<<<grid, 128, 200 * 1024>>>(c, cyc);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      best = std::min(best, ms);
      std::vector<long long> h(grid);
      CK(cudaMemcpy(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost));
      long long m = 0;
      for (auto v : h) m = std::max(m, v);
      cy = std::min(cy, m);
    }
    const double nm = double(c.iters) * per_commit;
    const double flops = nm * 2.0 * 128 * n * 16 * grid;
    printf("%-38s N=%3d chunk=%3d SBOrows=%2d per_commit=%3d nacc=%d grid=%3d : %6.1f cyc/MMA (ideal %5.1f)  %7.1f TFLOP/s\n", name, n,
           chunk, sbo_rows, per_commit, nacc, grid, cy / nm, n / 2.0, flops / (best * 1e-3) / 1e12);
  }
  cudaFree(cyc);
}


// minimal-instruction variant: 4 fixed descriptor pairs held in registers, `per_commit` MMAs per commit
__global__ void __launch_bounds__(128, 1) mma_min_kernel(MmaCfg c, long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t done[2];
  __shared__ __align__(8) uint64_t scratch[2];
  __shared__ __align__(8) uint64_t fin;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&done[0], 1);
    mbar_init(&done[1], 1);
    mbar_init(&scratch[0], 1);
    mbar_init(&scratch[1], 1);
    mbar_init(&fin, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (warp == 1 && elect_one()) {
    const uint32_t layout = layout_type_for_swizzle(c.chunk_bytes);
    const uint32_t idesc = make_idesc(1, 128, c.n, 0, 0);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 64 * 1024;
    uint64_t ad[4], bd[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ad[k] = make_smem_desc(a0 + k * 32, 16, 8 * c.chunk_bytes, layout);
      bd[k] = make_smem_desc(b0 + k * 32, 16, 8 * c.chunk_bytes, layout);
    }
    const int groups = c.per_commit / 4;
    long long t0 = clock64();
    for (int it = 0; it < c.iters; ++it) {
      const uint32_t d = tmem + (it % c.nacc) * 256;
      for (int g = 0; g < groups; ++g) {
        umma_ss<false>(d, ad[0], bd[0], idesc, 1);
        umma_ss<false>(d, ad[1], bd[1], idesc, 1);
        umma_ss<false>(d, ad[2], bd[2], idesc, 1);
        umma_ss<false>(d, ad[3], bd[3], idesc, 1);
      }
      // modes: 0 commit + wait(prev); 1 commit only (scratch barrier, never waited); 2 no commit; 4 two commits, no wait
      if (c.wait_each == 0) {
        umma_commit(&done[it & 1]);
        if (it >= 1) mbar_wait(&done[(it - 1) & 1], ((it - 1) >> 1) & 1);
      } else if (c.wait_each == 1) {
        umma_commit(&scratch[0]);
      } else if (c.wait_each == 4) {
        umma_commit(&scratch[0]);
        umma_commit(&scratch[1]);
      }
    }
    if (c.wait_each == 0) {
      mbar_wait(&done[(c.iters - 1) & 1], ((c.iters - 1) >> 1) & 1);
    } else {
      umma_commit(&fin);
      mbar_wait(&fin, 0);
    }
    long long t1 = clock64();
    cycles_out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem, 512);
  }
}
static void run_mma_min(int n, int per_commit, int nacc, int mode = 0) {
  MmaCfg c{};
  c.n = n, c.chunk_bytes = 128, c.per_commit = per_commit, c.iters = 7200 / per_commit, c.nacc = nacc;
  c.wait_each = mode;
  long long* cyc;
  CK(cudaMalloc(&cyc, 148 * 8));
  CK(cudaFuncSetAttribute(mma_min_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  long long cy = 1LL << 60;
  for (int rep = 0; rep < 3; ++rep) {
    mma_min_kernel<<<148, 128, 200 * 1024>>>(c, cyc);
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(148);
    CK(cudaMemcpy(h.data(), cyc, 148 * 8, cudaMemcpyDeviceToHost));
    long long m = 0;
    for (auto v : h) m = std::max(m, v);
    cy = std::min(cy, m);
  }
  const double nm = double(c.iters) * per_commit;
  printf("minimal issue loop N=%3d per_commit=%3d nacc=%d mode=%d (0 commit+wait prev, 1 commit only, 2 no commit, 3 commit+raw try_wait) : %6.1f cyc/MMA (ideal %5.1f)\n", n, per_commit, nacc, mode, cy / nm, n / 2.0);
  cudaFree(cyc);
}


// =====================================================================================
// tcgen05.mma with the A operand in TENSOR MEMORY (wgrad v2 design question): A [M=128][K] bf16 is written by the four
// warps with tcgen05.st (lane = row m, 32-bit column j holds K elements 2j, 2j+1), B [K pixels][N channels] is the
// MN-major SW128 tile a TMA box of an NHWC tensor gives (as in vk_wgrad.cuh).  Checks D = A * B^T-style product
// against the host, then times M=128 x N x K=16 MMAs with A from TMEM vs A from shared memory.
// =====================================================================================
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

struct TsCfg {
  int shift;      // correctness mode: the B view starts `shift` pixel rows into the loaded tile (unaligned to the 8-row atom)
  int n;          // GEMM N (channels of B)
  int k;          // K (pixel rows), multiple of 16, <= 128
  int iters;      // timing loop: MMAs issued
  int mode;       // 0: correctness (TS); 1: timing TS; 2: timing SS (A MN-major from smem)
};

__global__ void __launch_bounds__(128, 1) tsmma_kernel(const __grid_constant__ CUtensorMap tmb, const uint16_t* __restrict__ a_g,
                                                       TsCfg c, float* out, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full, done;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int blk_bytes = (c.k + 24) * 128;          // one 64-channel block: K + 24 rows of 128 B (room for shifted views)
  const int n_blocks = (c.n + 63) / 64;
  if (threadIdx.x == 0) {
    mbar_init(&full, 1);
    mbar_init(&done, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t tmem_a = tmem + 256;              // A: columns [256, 256 + k/2)
  // ---- A -> TMEM: lane m = row m, column j = (A[m][2j], A[m][2j+1]) ----
  {
    const int m = threadIdx.x;
    const uint32_t* row = reinterpret_cast<const uint32_t*>(a_g + size_t(m) * c.k);
    for (int j0 = 0; j0 < c.k / 2; j0 += 16) {
      uint32_t r[16];
      for (int i = 0; i < 16; ++i) r[i] = row[j0 + i];
      tmem_st16(tmem_a + (uint32_t(warp * 32) << 16) + j0, r);
    }
    tmem_st_wait();
  }
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&full, n_blocks * blk_bytes);
    for (int j = 0; j < n_blocks; ++j) tma_load_3d(smem + j * blk_bytes, &tmb, &full, j * 64, 0, 0);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (threadIdx.x == 0) {
    mbar_wait(&full, 0);
    tc_fence_after_sync();
    // B: MN-major SW128 (layout 2), LBO = block pitch, SBO = 1024 (8 pixel rows); A from TMEM is K-major by construction
    const uint32_t idesc_ts = make_idesc(1, 128, c.n, 0, 1);
    const uint32_t idesc_ss = make_idesc(1, 128, c.n, 1, 1);
    const uint32_t b0 = smem_u32(smem);
    const int ksteps = c.k / 16;
    if (c.mode == 0) {
      for (int kk = 0; kk < ksteps; ++kk) {
        const uint64_t bd = make_smem_desc(b0 + c.shift * 128 + kk * 2048, blk_bytes, 1024, 2);
        umma_ts(tmem, tmem_a + kk * 8, bd, idesc_ts, kk > 0);
      }
      umma_commit(&done);
    } else {
      // descriptors precomputed, issue loop fully unrolled (a run-time loop is issue-bound at ~140 clk per MMA)
      uint64_t bd[8], ad[8];
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        bd[kk] = make_smem_desc(b0 + kk * 2048, blk_bytes, 1024, 2);
        ad[kk] = make_smem_desc(b0 + kk * 2048, blk_bytes, 1024, 2);     // any resident tile: timing only
      }
      const long long t0 = clock64();
      if (c.mode == 1) {
        for (int it = 0; it < c.iters; it += 8) {
          const uint32_t d = tmem + ((it >> 3) & 1) * 128;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_ts(d, tmem_a + kk * 8, bd[kk], idesc_ts, 1);
        }
      } else {
        for (int it = 0; it < c.iters; it += 8) {
          const uint32_t d = tmem + ((it >> 3) & 1) * 128;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_ss<false>(d, ad[kk], bd[kk], idesc_ss, 1);
        }
      }
      umma_commit(&done);
      mbar_wait(&done, 0);
      cyc[0] = clock64() - t0;
    }
  }
  mbar_wait(&done, 0);
  tc_fence_after_sync();
  if (c.mode == 0) {
    for (int jc = 0; jc < c.n; jc += 16) {
      uint32_t rr[16];
      __syncwarp();
      tmem_ld16(tmem + (uint32_t(warp * 32) << 16) + jc, rr);
      tmem_ld_wait();
      for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * c.n + jc + i] = __uint_as_float(rr[i]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem, 512);
  }
}

static void run_tsmma() {
  const int K = 128;
  for (int N : {96, 64, 192}) {
    const int KB = K + 24;
    std::vector<float> a(size_t(128) * K), b(size_t(KB) * N);
    std::vector<uint16_t> a16(a.size()), b16(b.size());
    srand(7);
    for (size_t i = 0; i < a.size(); ++i) a[i] = bf16_round((rand() % 2001 - 1000) / 1000.f), a16[i] = bf16_bits(a[i]);
    for (size_t i = 0; i < b.size(); ++i) b[i] = bf16_round((rand() % 2001 - 1000) / 1000.f), b16[i] = bf16_bits(b[i]);
    void *da, *db;
    float* dout;
    long long* dcyc;
    CK(cudaMalloc(&da, a16.size() * 2));
    CK(cudaMalloc(&db, b16.size() * 2));
    CK(cudaMalloc(&dout, 128 * N * 4));
    CK(cudaMalloc(&dcyc, 8));
    CK(cudaMemcpy(da, a16.data(), a16.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, b16.data(), b16.size() * 2, cudaMemcpyHostToDevice));
    uint64_t dims[3] = {uint64_t(N), uint64_t(KB), 1};
    uint64_t strides[2] = {uint64_t(N) * 2, uint64_t(N) * 2 * KB};
    uint32_t box[3] = {64, uint32_t(KB), 1};
    CUtensorMap tmb = make_map(db, 3, dims, strides, box, 128);
    CK(cudaFuncSetAttribute(tsmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    for (int shift : {0, 1, 2, 3, 8, 17, 18, 19}) {
    TsCfg c{shift, N, K, 0, 0};
    CK(cudaMemset(dout, 0, 128 * N * 4));
    tsmma_kernel<<<1, 128, 100 * 1024>>>(tmb, reinterpret_cast<const uint16_t*>(da), c, dout, dcyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("tsmma N=%d: CUDA error %s\n", N, cudaGetErrorString(e));
      exit(1);
    }
    std::vector<float> h(128 * N);
    CK(cudaMemcpy(h.data(), dout, h.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        double acc = 0;
        for (int k = 0; k < K; ++k) acc += double(a[size_t(m) * K + k]) * b[size_t(k + shift) * N + n];
        maxerr = std::max(maxerr, std::fabs(acc - h[m * N + n]));
      }
    printf("tsmma (A in TMEM, B MN-major SW128 view shifted by %2d pixel rows) M=128 N=%3d K=%d : max|err| = %.3e %s\n", shift, N, K,
           maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
    }
    for (int mode : {1, 2}) {
      TsCfg t{0, N, K, 4096, mode};
      tsmma_kernel<<<1, 128, 100 * 1024>>>(tmb, reinterpret_cast<const uint16_t*>(da), t, dout, dcyc);
      CK(cudaDeviceSynchronize());
      long long cy;
      CK(cudaMemcpy(&cy, dcyc, 8, cudaMemcpyDeviceToHost));
      printf("  rate N=%3d %s : %6.1f cyc/MMA (ideal %5.1f)\n", N, mode == 1 ? "A from TMEM  " : "A from smem  ",
             double(cy) / 4096, N / 2.0);
    }
    cudaFree(da), cudaFree(db), cudaFree(dout), cudaFree(dcyc);
  }
}


// =====================================================================================
// Register <-> (lane, column) mapping of tcgen05.st.16x256b.x1 (4 registers per thread), read back with the
// 32x32b shape whose mapping is known (thread = lane, register = column).
// =====================================================================================
__global__ void __launch_bounds__(128, 1) stmap_kernel(uint32_t* out) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 32);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  // every warp writes its lane quarter: two 16-lane halves, value = (warp << 24) | (half << 20) | (lane << 8) | reg
  for (int hf = 0; hf < 2; ++hf) {
    uint32_t r[4];
    for (int i = 0; i < 4; ++i) r[i] = (uint32_t(warp) << 24) | (uint32_t(hf) << 20) | (uint32_t(lane) << 8) | uint32_t(i);
    const uint32_t taddr = tmem + (uint32_t(warp * 32 + hf * 16) << 16);
    asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3])
                 : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t v[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(tmem + (uint32_t(warp * 32) << 16)));
  tmem_ld_wait();
  for (int i = 0; i < 8; ++i) out[threadIdx.x * 8 + i] = v[i];
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem, 32);
  }
}

static void run_stmap() {
  uint32_t* d;
  CK(cudaMalloc(&d, 128 * 8 * 4));
  stmap_kernel<<<1, 128>>>(d);
  CK(cudaDeviceSynchronize());
  std::vector<uint32_t> h(128 * 8);
  CK(cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost));
  printf("tcgen05.st.16x256b.x1: TMEM (lane, column) <- (warp, half, thread, register)\n");
  for (int l = 0; l < 40; ++l) {
    printf("lane %3d:", l);
    for (int c = 0; c < 8; ++c) {
      const uint32_t v = h[l * 8 + c];
      printf("  w%u h%u t%2u r%u", v >> 24, (v >> 20) & 15, (v >> 8) & 255, v & 255);
    }
    printf("\n");
  }
  cudaFree(d);
}

int main(int argc, char** argv) {
  const char* what = argc > 1 ? argv[1] : "all";
  if (!strcmp(what, "tsmma")) run_tsmma();
  if (!strcmp(what, "stmap")) run_stmap();
  if (!strcmp(what, "shift") || !strcmp(what, "all")) run_shift();
  if (!strcmp(what, "tma") || !strcmp(what, "all")) {
    // L2-resident (16 images of 128x128xC bf16: 50 MB at C=96) and HBM-streaming (64 images at C=192: 400 MB)
    run_tma("C96 inner64B SW64 box16x10", 96, 32, 64, 16, 10, 3, 4, 16);
    run_tma("C96 inner64B SW64 box16x10", 96, 32, 64, 16, 10, 3, 8, 16);
    run_tma("C96 inner64B SW64 box16x10", 96, 32, 64, 16, 10, 3, 16, 16);
    run_tma("C96 inner64B SW64 box10x18 (halo slab)", 96, 32, 64, 10, 18, 3, 8, 16);
    run_tma("C96 inner128B SW128 ch0-63 box16x10", 96, 64, 128, 16, 10, 1, 8, 16);
    run_tma("C96 inner128B SW128 ch0-63 box10x18", 96, 64, 128, 10, 18, 1, 8, 16);
    run_tma("C96 inner192B noswizzle box16x10", 96, 96, 0, 16, 10, 1, 4, 16);
    run_tma("C192 inner128B SW128 box16x10", 192, 64, 128, 16, 10, 3, 4, 8);
    run_tma("C192 inner128B SW128 box16x10", 192, 64, 128, 16, 10, 3, 8, 8);
    run_tma("C192 inner128B SW128 box10x18", 192, 64, 128, 10, 18, 3, 8, 8);
    run_tma("C192 inner64B SW64 box16x10", 192, 32, 64, 16, 10, 6, 8, 8);
    run_tma("C192 inner128B SW128 box16x10 HBM(400MB)", 192, 64, 128, 16, 10, 3, 8, 64);
    run_tma("C64 inner128B SW128 box16x10", 64, 64, 128, 16, 10, 1, 8, 16);
    run_tma("C192 inner128B SW128 box32x8 (256 rows)", 192, 64, 128, 32, 8, 3, 6, 8);
  }
  if (!strcmp(what, "mma2")) {
    for (int n : {32, 96, 128, 256})
      for (int pc : {4, 12, 36, 144}) run_mma_min(n, pc, 2);
    run_mma_min(96, 36, 1);
  }
  if (!strcmp(what, "mma3")) {
    for (int n : {96, 256})
      for (int mode : {2, 1, 4, 0})
        for (int pc : {4, 12, 36}) run_mma_min(n, pc, 2, mode);
  }
  if (!strcmp(what, "mma")) {
    for (int n : {16, 32, 64, 96, 128, 160, 192, 224, 256}) run_mma("SW128 dense A", n, 128, 8, 36, 2);
    for (int n : {64, 96, 192}) run_mma("SW128 halo-slab A (SBO 18 rows)", n, 128, 18, 36, 2);
    for (int n : {64, 96, 192}) run_mma("SW64 dense A", n, 64, 8, 36, 2);
    for (int n : {96}) run_mma("SW64 halo-slab A", n, 64, 18, 36, 2);
    for (int n : {96, 192}) run_mma("SW128 dense, 1 accumulator", n, 128, 8, 36, 1);
    for (int n : {96}) run_mma("SW128 dense, commit every 12", n, 128, 8, 12, 2);
    for (int n : {96}) run_mma("SW128 dense, commit+wait every 12", n, 128, 8, 12, 2, 1);
    for (int n : {96}) run_mma("SW128 dense, commit+wait every 36", n, 128, 8, 36, 2, 1);
  }
  if (!strcmp(what, "tma5")) {
    for (int iss : {1, 2, 4, 8}) run_tma_multi("4D 16x8 warps", 0, iss, 12 / iss > 4 ? 4 : (12 / iss ? 12 / iss : 1), 16, 8, 0);
    for (int iss : {1, 2, 4, 8}) run_tma_multi("4D 16x8 lanes of one warp", 1, iss, 1, 16, 8, 0);
    for (int iss : {4}) run_tma_multi("4D 18x18 warps", 0, iss, 1, 18, 18, 0);
    for (int iss : {1, 2, 4}) run_tma_multi("1D bulk 12KB", 2, iss, 4, 0 + 16, 8, 12288);
    for (int iss : {1, 2}) run_tma_multi("1D bulk 24KB", 2, iss, 4, 16, 8, 24576);
    for (int iss : {1, 2}) run_tma_multi("1D bulk 6KB", 2, iss, 8, 16, 8, 6144);
  }
  if (!strcmp(what, "tma4")) {
    const int shapes[6][2] = {{128, 1}, {64, 2}, {32, 4}, {16, 8}, {8, 16}, {4, 32}};
    for (int rank : {3, 4})
      for (auto& sh : shapes) run_tma_batch("C192 SW128 batch8 by shape", rank, 192, 64, 128, sh[0], sh[1], 8, 1, 8);
    for (auto& sh : shapes) run_tma_batch("C192 SW128 batch1 by shape", 4, 192, 64, 128, sh[0], sh[1], 1, 1, 8);
    run_tma_batch("C192 SW128 2D batch8", 2, 192, 64, 128, 16, 8, 8, 1, 8);
    run_tma_batch("C96 SW64 4D 16x8 batch8", 4, 96, 32, 64, 16, 8, 8, 1, 16);
    run_tma_batch("C96 SW64 2D batch8", 2, 96, 32, 64, 16, 8, 8, 1, 16);
    run_tma_batch("C192 SW128 4D 16x8 batch8 2 issuers", 4, 192, 64, 128, 16, 8, 4, 2, 8);
  }
  if (!strcmp(what, "tma3")) {
    for (int rank : {2, 3, 4}) run_tma_ring("C192 SW128", rank, 192, 64, 128, 16, 8, 8, 0, 8);
    for (int rank : {3, 4}) run_tma_ring("C192 SW128 halo", rank, 192, 64, 128, 16, 8, 8, 1, 8);
    for (int rank : {2, 3, 4}) run_tma_ring("C192 SW128 12 stages", rank, 192, 64, 128, 16, 8, 12, 0, 8);
    for (int rank : {3, 4}) run_tma_ring("C192 SW128 18x18 slab", rank, 192, 64, 128, 18, 18, 4, 1, 8);
    for (int rank : {3, 4}) run_tma_ring("C192 SW128 10x18 slab", rank, 192, 64, 128, 10, 18, 8, 1, 8);
    for (int rank : {3, 4}) run_tma_ring("C192 SW128 128x1 line", rank, 192, 64, 128, 128, 1, 8, 0, 8);
    for (int rank : {3, 4}) run_tma_ring("C192 SW128 64x2", rank, 192, 64, 128, 64, 2, 8, 0, 8);
    for (int rank : {3, 4}) run_tma_ring("C192 SW128 32x4", rank, 192, 64, 128, 32, 4, 8, 0, 8);
    for (int rank : {3, 4}) run_tma_ring("C192 SW128 8x16", rank, 192, 64, 128, 8, 16, 8, 0, 8);
    for (int rank : {2, 3, 4}) run_tma_ring("C64 SW128 (dense)", rank, 64, 64, 128, 16, 8, 8, 0, 16);
    for (int rank : {2, 4}) run_tma_ring("C96 SW64", rank, 96, 32, 64, 16, 8, 12, 0, 16);
  }
  if (!strcmp(what, "tma2")) {
    for (int batch : {1, 2, 4, 8}) run_tma_batch("C192 2D SW128 128rows", 2, 192, 64, 128, 16, 8, batch, 1, 8);
    for (int batch : {1, 4}) run_tma_batch("C192 4D SW128 16x8", 4, 192, 64, 128, 16, 8, batch, 1, 8);
    for (int batch : {1, 4}) run_tma_batch("C192 2D SW128 32rows", 2, 192, 64, 128, 8, 4, batch, 1, 8);
    for (int batch : {1, 4}) run_tma_batch("C192 2D SW128 256rows", 2, 192, 64, 128, 32, 8, batch, 1, 8);
    for (int batch : {1, 4}) run_tma_batch("C96 2D SW64 128rows", 2, 96, 32, 64, 16, 8, batch, 1, 16);
    for (int batch : {1, 4}) run_tma_batch("C192 2D SW128 128rows 2 issuers", 2, 192, 64, 128, 16, 8, batch, 2, 8);
    for (int batch : {1, 4}) run_tma_batch("C192 2D SW128 128rows 4 issuers", 2, 192, 64, 128, 16, 8, batch, 4, 8);
    for (int batch : {4}) run_tma_batch("C64 2D SW128 128rows (dense rows)", 2, 64, 64, 128, 16, 8, batch, 1, 16);
  }
  return 0;
}
