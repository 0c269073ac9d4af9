// umma_pair_probe.cu -- hardware probe: tcgen05.mma.cta_group::2 (a CTA pair on one TPC executing ONE M = 256
// instruction) with operands placed by ordinary shared-memory stores, as the gather kernel places them.
//
// Questions: (1) with M = 256, N = 256 and both CTAs holding 128 rows of A at the SAME shared-memory offset, does
// CTA r's TMEM receive rows [128 r, 128 r + 128) x all 256 columns?  (2) B is split along N: does CTA r's
// shared memory hold B rows (= output columns) [128 r, 128 r + 128)?  (3) does the multicast commit
// (tcgen05.commit.cta_group::2 ... multicast::cluster, mask 0b11) release a barrier in both CTAs?
// Prints MATCH / MISMATCH against an integer reference (exact in fp32) for each hypothesis about the B split.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O2 -o tools/umma_pair_probe.bin tools/umma_pair_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../cbinfer_b200/csrc/conv_umma.cuh"

namespace cb {
thread_local char g_err[512] = "";
int fail(int code, const char* fmt, ...) { (void)fmt; return code; }
int sm_count() { return 148; }
bool pdl_enabled() { return false; }
}  // namespace cb

using namespace cb;

constexpr int K = 64, M2 = 256, N2 = 256;

__device__ __forceinline__ void umma_pair_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

struct PairCtrl {
  uint64_t bar_mma;
  uint32_t tmem_base, pad;
};

// A: [256][64] bf16 (row-major), B: [256][64] bf16 (B[n][k]); out: [256][256] fp32 = A * B^T
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
pair_probe_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  uint8_t* a_tile = smem;                       // 128 rows x 128 B, SWIZZLE_128B K-major
  uint8_t* b_tile = smem + 128 * 128;           // 128 rows (N half) x 128 B
  PairCtrl* ctrl = reinterpret_cast<PairCtrl*>(smem + 2 * 128 * 128);
  for (int i = tid; i < 128 * K; i += blockDim.x) {
    const int r = i / K, k = i - r * K;
    const uint32_t off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 3) ^ (r & 7))) << 4) + (k & 7) * 2);
    *reinterpret_cast<__nv_bfloat16*>(a_tile + off) = A[(rank * 128 + r) * K + k];
    *reinterpret_cast<__nv_bfloat16*>(b_tile + off) = B[(rank * 128 + r) * K + k];
  }
  if (tid == 0) {
    mbar_init(&ctrl->bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&ctrl->tmem_base)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                           // both CTAs' operands, barriers and TMEM exist
  tc_fence_after();
  const uint32_t tmem = ctrl->tmem_base;
  if (rank == 0 && tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N2 >> 3) << 17) | ((uint32_t)(M2 >> 4) << 24);
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t ad = umma_desc(smem_u32(a_tile) + ks * 32), bd = umma_desc(smem_u32(b_tile) + ks * 32);
      umma_pair_f16(tmem, ad, bd, idesc, ks ? 1u : 0u);
    }
    umma_commit_pair(&ctrl->bar_mma);
  }
  mbar_wait(&ctrl->bar_mma, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N2; c0 += 16) {
    uint32_t acc[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, acc);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[(size_t)(rank * 128 + row) * N2 + c0 + i] = __uint_as_float(acc[i]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                           // nobody frees TMEM while the peer still reads
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

int main() {
  std::vector<__nv_bfloat16> hA(M2 * K), hB(N2 * K);
  std::vector<float> fA(M2 * K), fB(N2 * K);
  srand(1);
  for (int i = 0; i < M2 * K; ++i) { fA[i] = (float)(rand() % 7 - 3); hA[i] = __float2bfloat16(fA[i]); }
  for (int i = 0; i < N2 * K; ++i) { fB[i] = (float)(rand() % 5 - 2); hB[i] = __float2bfloat16(fB[i]); }
  __nv_bfloat16 *dA, *dB;
  float* dO;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dO, (size_t)M2 * N2 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dO, 0xff, (size_t)M2 * N2 * 4);
  const size_t smem = 2 * 128 * 128 + 1024;
  cudaFuncSetAttribute(pair_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  pair_probe_kernel<<<2, 128, smem>>>(dA, dB, dO);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> hO((size_t)M2 * N2);
  cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
  // hypothesis 0: out[m][n] = sum_k A[m][k] B[n][k] (B rows [128 r, 128 r + 128) live in CTA r)
  // hypothesis 1: the halves of B are swapped between the CTAs
  for (int hyp = 0; hyp < 2; ++hyp) {
    long bad = 0;
    for (int m = 0; m < M2; ++m)
      for (int n = 0; n < N2; ++n) {
        const int nn = hyp ? (n + 128) % 256 : n;
        float s = 0.f;
        for (int k = 0; k < K; ++k) s += fA[m * K + k] * fB[nn * K + k];
        if (s != hO[(size_t)m * N2 + n]) ++bad;
      }
    printf("hypothesis %d (%s): %s (%ld of %d wrong)\n", hyp, hyp ? "B halves swapped" : "CTA r holds B rows 128r..128r+127",
           bad ? "MISMATCH" : "MATCH", bad, M2 * N2);
  }
  printf("sample out[0][0..3] = %g %g %g %g, out[128][128..131] = %g %g %g %g\n", hO[0], hO[1], hO[2], hO[3],
         hO[128 * 256 + 128], hO[128 * 256 + 129], hO[128 * 256 + 130], hO[128 * 256 + 131]);
  return 0;
}
