// Experiment: tcgen05.mma with the A operand resident in tensor memory (written once with tcgen05.st).
//   1. numerics: D[128 x 64] = A[128 x 64] * B[64 x 64]^T with small-integer operands against a host reference,
//      A rows <-> TMEM lanes, two consecutive k elements per 32-bit column (low half = lower k)
//   2. cost per MMA instruction (M=128, K=16): A from shared memory vs A from tensor memory, N = 64 / 128 / 256
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/tmem_a_test tools/tmem_a_test.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../parrot_b200/csrc/ptx.cuh"
using namespace pb;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

// smem byte offset of element (row, k) of a K-major 128B-swizzled bf16 tile (rows of 64 k)
__device__ __host__ inline int sw128_off(int row, int k) {
  const int chunk = (k * 2) >> 4, within = (k * 2) & 15;
  return (row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4) + within;
}

struct TArgs { const __nv_bfloat16* A; const __nv_bfloat16* B; float* D; long long* cyc; int N; };

__global__ void __launch_bounds__(160, 1) test_kernel(const TArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;             // [128][64] swizzled (for the SS reference timing)
  uint8_t* sB = smem + 16384;     // [256][64] swizzled
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(bar, 1); for (int i = 2; i < 6; ++i) mbar_init(bar + i, 1); fence_barrier_init(); }
  if (warp == 4) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  // operands -> smem (swizzled)
  for (int e = tid; e < 128 * 64; e += blockDim.x) {
    const int r = e >> 6, k = e & 63;
    *reinterpret_cast<__nv_bfloat16*>(sA + sw128_off(r, k)) = a.A[e];
  }
  for (int e = tid; e < 256 * 64; e += blockDim.x) {
    const int r = e >> 6, k = e & 63;
    *reinterpret_cast<__nv_bfloat16*>(sB + sw128_off(r, k)) = a.B[(r & 63) * 64 + k];
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_slot;
  const uint32_t colA = 384;   // A region: 32 columns for one 64-wide k block
  if (warp < 4) {
    const int m = warp * 32 + lane;
    for (int c0 = 0; c0 < 32; c0 += 8) {
      uint32_t r[8];
      for (int j = 0; j < 8; ++j) {
        const int k = (c0 + j) * 2;
        const uint16_t lo = *reinterpret_cast<const uint16_t*>(&a.A[m * 64 + k]);
        const uint16_t hi = *reinterpret_cast<const uint16_t*>(&a.A[m * 64 + k + 1]);
        r[j] = (uint32_t)lo | ((uint32_t)hi << 16);
      }
      tmem_st_32x8(tb + ((uint32_t)(warp * 32) << 16) + colA + c0, r);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t idesc = umma_idesc_bf16(128, a.N);
  const uint64_t db = umma_desc_sw128(sB), da = umma_desc_sw128(sA);
  uint32_t ph = 0;
  if (tid == 128) {
    // ---- numerics: D (cols 0..N) = A_tmem * B
    for (int k = 0; k < 4; ++k) umma_bf16_ts(tb, tb + colA + k * 8, db + (uint64_t)((k * 32) >> 4), idesc, k != 0);
    umma_commit(bar);
    mbar_wait(bar, ph); ph ^= 1;
    tc_fence_after();
  }
  __syncthreads();
  tc_fence_after();
  if (warp < 4) {
    for (int n0 = 0; n0 < 64; n0 += 8) {
      float v[8];
      tmem_ld_32x8(tb + ((uint32_t)(warp * 32) << 16) + n0, v);
      tmem_ld_wait();
      for (int j = 0; j < 8; ++j) a.D[(warp * 32 + lane) * 64 + n0 + j] = v[j];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 128) {
    // ---- timing: 2000 MMAs each, commit + wait at the end
    for (int mode = 0; mode < 2; ++mode) {
      const long long c0 = clock64();
      for (int i = 0; i < 2000; ++i) {
        const int k = i & 3;
        if (mode == 0) umma_bf16(tb, da + (uint64_t)((k * 32) >> 4), db + (uint64_t)((k * 32) >> 4), idesc, 1);
        else umma_bf16_ts(tb, tb + colA + k * 8, db + (uint64_t)((k * 32) >> 4), idesc, 1);
      }
      umma_commit(bar);
      mbar_wait(bar, ph); ph ^= 1;
      a.cyc[mode] = clock64() - c0;
    }
    // ---- pipeline handshake cost: groups of 12 MMAs, one commit per group on a ring of 4 barriers
    //   mode 2: commits only ; mode 3: + wait for the group committed 3 groups ago + fence::after_thread_sync ;
    //   mode 4: as 3 without the fence ; mode 5: as 3 with A in TMEM
    uint64_t* ring = bar + 2;
    for (int mode = 2; mode < 6; ++mode) {
      uint32_t rph[4] = {0, 0, 0, 0};
      const long long c0 = clock64();
      const int groups = 168;
      for (int g = 0; g < groups; ++g) {
        if (mode >= 3 && g >= 3) {
          const int w = (g - 3) & 3;
          mbar_wait(&ring[w], rph[w]); rph[w] ^= 1;
          if (mode != 4) tc_fence_after();
        }
        for (int i = 0; i < 12; ++i) {
          const int k = i & 3;
          if (mode == 5) umma_bf16_ts(tb, tb + colA + k * 8, db + (uint64_t)((k * 32) >> 4), idesc, 1);
          else umma_bf16(tb, da + (uint64_t)((k * 32) >> 4), db + (uint64_t)((k * 32) >> 4), idesc, 1);
        }
        umma_commit(&ring[g & 3]);
      }
      // drain
      for (int g = (mode >= 3 ? groups - 3 : 0); g < groups; ++g) {
        if (mode < 3 && g < groups - 4) { continue; }
      }
      umma_commit(bar);
      mbar_wait(bar, ph); ph ^= 1;
      a.cyc[mode] = clock64() - c0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tb, 512);
}

int main() {
  CK(cudaSetDevice(0));
  std::vector<__nv_bfloat16> hA(128 * 64), hB(64 * 64);
  std::vector<float> fA(128 * 64), fB(64 * 64);
  srand(1);
  for (int i = 0; i < 128 * 64; ++i) { fA[i] = (float)((rand() % 15) - 7); hA[i] = __float2bfloat16(fA[i]); }
  for (int i = 0; i < 64 * 64; ++i) { fB[i] = (float)((rand() % 15) - 7); hB[i] = __float2bfloat16(fB[i]); }
  __nv_bfloat16 *dA, *dB; float* dD; long long* dc;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, 128 * 64 * 4)); CK(cudaMalloc(&dc, 64));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  for (int N : {64, 128, 256}) {
    TArgs a{dA, dB, dD, dc, N};
    test_kernel<<<1, 160, 64 * 1024>>>(a);
    CK(cudaDeviceSynchronize());
    std::vector<float> hD(128 * 64);
    long long cyc[8];
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cyc, dc, 48, cudaMemcpyDeviceToHost));
    int bad = 0; double maxerr = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 64; ++n) {
        float ref = 0;
        for (int k = 0; k < 64; ++k) ref += fA[m * 64 + k] * fB[n * 64 + k];
        const double e = fabs(ref - hD[m * 64 + n]);
        if (e > 1e-3) { if (bad < 5) printf("  mismatch m=%d n=%d ref=%g got=%g\n", m, n, ref, hD[m * 64 + n]); ++bad; }
        if (e > maxerr) maxerr = e;
      }
    printf("N=%3d  A-in-TMEM numerics: %s (bad %d, max err %g) ; cycles per MMA: A in smem %.1f, A in TMEM %.1f\n", N,
           bad ? "MISMATCH" : "exact", bad, maxerr, cyc[0] / 2000.0, cyc[1] / 2000.0);
    printf("        per MMA with one commit per 12: commits only %.1f ; + ring wait + fence %.1f ; + ring wait, no fence %.1f ; TS + ring wait + fence %.1f\n",
           cyc[2] / 2016.0, cyc[3] / 2016.0, cyc[4] / 2016.0, cyc[5] / 2016.0);
  }
  return 0;
}
