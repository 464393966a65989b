// Micro-benchmark: per-SM throughput of the operand-staging paths (L2 -> shared memory) used by the job engine.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/tma_bw tools/tma_bw.cu -lcuda
//   build/tma_bw
// One producer thread per CTA keeps S stages in flight; a consumer thread frees a stage as soon as it lands.
//   mode 0: 3-D tensor map (k, row, plane) box {64,128,2} = 32 KB, tile-contiguous source (the engine's weight path)
//   mode 1: two 2-D boxes {64,128} of 16 KB per stage (hi, lo issued separately)
//   mode 2: two cp.async.bulk 1-D copies of 16 KB per stage (no tensor map)
//   mode 3: 3-D tensor map box {64,128,2} over row-major planes with a 2 KB row pitch (the activation path)
//   mode 4: one cp.async.bulk 1-D copy of 32 KB per stage
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../parrot_b200/csrc/ptx.cuh"
using namespace pb;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void bulk_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ unsigned long long gtime2() {
  unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t;
}

struct Args {
  const CUtensorMap* maps;   // [0] 3-D tiled, [1] 2-D tiled hi, [2] 2-D tiled lo, [3] 3-D strided
  const uint8_t* base;       // tiled planes: hi at base, lo at base + plane_bytes
  size_t plane_bytes;
  int mode, stages, boxes_per_cta, reps, consume_cycles, ncols;
  unsigned long long* out;   // [cta][2]
};

__global__ void __launch_bounds__(64, 1) bw_kernel(const Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)a.stages * 32768);
  uint64_t* empty = full + 16;
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  __shared__ uint32_t tmem_slot;
  if (threadIdx.x < 32) { tmem_alloc(&tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int total = a.boxes_per_cta * a.reps;
  unsigned long long t0 = 0;
  if (threadIdx.x == 0) {
    t0 = gtime2();
    int st = 0; uint32_t ph = 0;
    for (int i = 0; i < total; ++i) {
      const int box = (int)blockIdx.x * a.boxes_per_cta + (i % a.boxes_per_cta);
      mbar_wait(&empty[st], ph ^ 1);
      uint8_t* dst = smem + (size_t)st * 32768;
      mbar_expect_tx(&full[st], 32768);
      if (a.mode == 0) tma_load_3d(dst, a.maps + 0, &full[st], 0, box * 128, 0);
      else if (a.mode == 1) {
        tma_load_2d(dst, a.maps + 1, &full[st], 0, box * 128);
        tma_load_2d(dst + 16384, a.maps + 2, &full[st], 0, box * 128);
      } else if (a.mode == 2) {
        bulk_1d(dst, a.base + (size_t)box * 16384, 16384, &full[st]);
        bulk_1d(dst + 16384, a.base + a.plane_bytes + (size_t)box * 16384, 16384, &full[st]);
      } else if (a.mode == 3) {
        // strided planes [rows][1024]: box (row block, k block)
        tma_load_3d(dst, a.maps + 3, &full[st], (box & 15) * 64, (box >> 4) * 128, 0);
      } else {
        bulk_1d(dst, a.base + (size_t)box * 32768, 32768, &full[st]);
      }
      if (++st == a.stages) { st = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    int st = 0; uint32_t ph = 0;
    const int nmma = -a.consume_cycles;            // < 0: consume with |x| tcgen05.mma per stage (M=128, N=ncols, K=16)
    const uint32_t idesc = umma_idesc_bf16(128, a.ncols);
    for (int i = 0; i < total; ++i) {
      mbar_wait(&full[st], ph);
      if (a.consume_cycles > 0) { const long long c0 = clock64(); while (clock64() - c0 < a.consume_cycles) {} }
      if (a.consume_cycles < 0) {
        tc_fence_after();
        const uint8_t* stp = smem + (size_t)st * 32768;
        const uint64_t da = umma_desc_sw128(stp), db = umma_desc_sw128(stp + 16384);
        for (int k = 0; k < nmma; ++k) {
          const uint64_t adv = (uint64_t)(((k & 3) * 32) >> 4);
          umma_bf16(tmem_base, da + adv, db + adv, idesc, (i | k) != 0);
        }
        umma_commit(&empty[st]);
      } else {
        mbar_arrive(&empty[st]);
      }
      if (++st == a.stages) { st = 0; ph ^= 1; }
    }
    if (a.consume_cycles < 0) {   // drain: wait until the last commit has arrived
      const int last = (total - 1) % a.stages;
      const uint32_t lph = (uint32_t)(((total - 1) / a.stages) & 1);
      mbar_wait(&empty[last], lph);
    }
    a.out[blockIdx.x * 2 + 1] = gtime2();
  }
  if (threadIdx.x == 0) a.out[blockIdx.x * 2] = t0;
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem_base, 256);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  CK(cudaSetDevice(0));
  const int max_boxes = 148 * 64;                       // 64 boxes of 32 KB per CTA max = 2 MB per CTA, 303 MB total
  const size_t plane_bytes = (size_t)max_boxes * 16384;
  uint8_t* buf;
  CK(cudaMalloc(&buf, 2 * plane_bytes));
  CK(cudaMemset(buf, 1, 2 * plane_bytes));
  CUtensorMap hm[4];
  {
    cuuint32_t es[4] = {1, 1, 1, 1};
    // tiled: [boxes*128][64]
    cuuint64_t d3[3] = {64, (cuuint64_t)max_boxes * 128, 2};
    cuuint64_t s3[2] = {128, plane_bytes};
    cuuint32_t b3[3] = {64, 128, 2};
    CUresult r = cuTensorMapEncodeTiled(&hm[0], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, d3, s3, b3, es,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode0 %d\n", r); return 1; }
    cuuint64_t d2[2] = {64, (cuuint64_t)max_boxes * 128};
    cuuint64_t s2[1] = {128};
    cuuint32_t b2[2] = {64, 128};
    for (int w = 0; w < 2; ++w) {
      r = cuTensorMapEncodeTiled(&hm[1 + w], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf + w * plane_bytes, d2, s2, b2, es,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) { printf("encode1 %d\n", r); return 1; }
    }
    // strided: [rows][1024] bf16, rows = max_boxes*128/16
    cuuint64_t d4[3] = {1024, (cuuint64_t)max_boxes * 128 / 16, 2};
    cuuint64_t s4[2] = {2048, plane_bytes};
    r = cuTensorMapEncodeTiled(&hm[3], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, d4, s4, b3, es,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode3 %d\n", r); return 1; }
  }
  CUtensorMap* dm;
  CK(cudaMalloc(&dm, sizeof hm));
  CK(cudaMemcpy(dm, hm, sizeof hm, cudaMemcpyHostToDevice));
  unsigned long long* out;
  CK(cudaMalloc(&out, 148 * 16));
  CK(cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  printf("mode stages grid boxes/cta(KB/cta) consume | median KB/us per SM | chip TB/s\n");
  const int modes[] = {0};
  for (int mode : modes)
    for (int stages : {2, 4, 6})
      for (int grid : {148})
        for (int boxes : {8})
         for (int ncols : {64, 128, 256})
          for (int cons : {0, -1, -4, -12, -24}) {
            if (cons == 0 && ncols != 64) continue;
            Args a;
            a.maps = dm; a.base = buf; a.plane_bytes = plane_bytes; a.mode = mode; a.stages = stages;
            a.boxes_per_cta = boxes; a.reps = 1920 / boxes; a.consume_cycles = cons; a.out = out; a.ncols = ncols;
            std::vector<unsigned long long> h(148 * 2);
            for (int it = 0; it < 2; ++it) {   // first run warms the L2
              bw_kernel<<<grid, 64, 200 * 1024>>>(a);
              CK(cudaDeviceSynchronize());
            }
            CK(cudaMemcpy(h.data(), out, grid * 16, cudaMemcpyDeviceToHost));
            std::vector<double> r;
            for (int c = 0; c < grid; ++c) r.push_back((double)(a.boxes_per_cta * a.reps) * 32.768 / ((double)(h[c * 2 + 1] - h[c * 2]) / 1e3));
            std::sort(r.begin(), r.end());
            const double med = r[r.size() / 2];
            printf("%d %d %3d %2d(%4d) N=%3d %4d | %7.1f | %.2f\n", mode, stages, grid, boxes, boxes * 32, ncols, cons, med,
                   med * grid / 1000.0);
          }
  return 0;
}
