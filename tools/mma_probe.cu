// Probe of tcgen05.mma kind::tf32 operand major-ness (debug aid, not part of the library).
// D[128][N] = A[128][K] * B[N][K]^T with K = 8*ksteps; operands staged as stacked [rows][128B] SWIZZLE_128B regions.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../bayesian-neural-network_b200/csrc/bbb_tc.cuh"
using namespace bbb::tc;

constexpr int REG = 16384;
struct Cfg { int a_mn, b_mn, N, ksteps; uint32_t lbo_a, sbo_a, lbo_b, sbo_b; };
// MN-major tf32: SWIZZLE_128B_BASE32B (layout type 1): 32-byte chunk index XOR (row & 3), atoms of 4 K rows
__device__ uint32_t mn32_off(int k, int mn31) { return (uint32_t)(k * 128 + ((((mn31 >> 3) ^ (k & 3)) & 3) << 5) + (mn31 & 7) * 4); }

__device__ uint64_t desc_generic(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t ltype = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1u << 46;
  d |= (uint64_t)ltype << 61;
  return d;
}

// A logical [128 m][K], B logical [N][K].  K-major tile: region r holds k in [32r, 32r+32), rows = m (or n).
// MN-major tile: region g holds m (or n) in [32g, 32g+32), rows = k.
__global__ void probe(const float *A, const float *B, float *D, Cfg c) {
  extern __shared__ uint8_t dsm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  uint8_t *base = (uint8_t *)(((uintptr_t)dsm + 1023) & ~(uintptr_t)1023);
  uint8_t *At = base, *Bt = base + 4 * REG;
  const int tid = threadIdx.x, K = 8 * c.ksteps;
  for (int i = tid; i < 8 * REG / 4; i += blockDim.x) ((float *)base)[i] = 0.f;
  __syncthreads();
  for (int idx = tid; idx < 128 * K; idx += blockDim.x) {
    int m = idx / K, k = idx % K;
    float v = A[m * K + k];
    if (c.a_mn) *(float *)(At + (m >> 5) * REG + mn32_off(k, m & 31)) = v;
    else *(float *)(At + (k >> 5) * REG + sw128_off(m, (k & 31) >> 2) + (k & 3) * 4) = v;
  }
  for (int idx = tid; idx < c.N * K; idx += blockDim.x) {
    int n = idx / K, k = idx % K;
    float v = B[n * K + k];
    if (c.b_mn) *(float *)(Bt + (n >> 5) * REG + mn32_off(k, n & 31)) = v;
    else *(float *)(Bt + (k >> 5) * REG + sw128_off(n, (k & 31) >> 2) + (k & 3) * 4) = v;
  }
  if (tid < 32) tmem_alloc(smem_u32(&tbase), 128);
  if (tid == 32) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tbase;
  if (tid == 0) {
    const uint32_t idesc = idesc_tf32_major(128, c.N, c.a_mn, c.b_mn);
    for (int ks = 0; ks < c.ksteps; ++ks) {
      uint64_t da = c.a_mn ? desc_generic(smem_u32(At) + ks * 1024, c.lbo_a, c.sbo_a, 1)
                           : desc_generic(smem_u32(At) + (ks >> 2) * REG + (ks & 3) * 32, c.lbo_a, c.sbo_a);
      uint64_t db = c.b_mn ? desc_generic(smem_u32(Bt) + ks * 1024, c.lbo_b, c.sbo_b, 1)
                           : desc_generic(smem_u32(Bt) + (ks >> 2) * REG + (ks & 3) * 32, c.lbo_b, c.sbo_b);
      mma_tf32(tmem, da, db, idesc, ks > 0);
    }
    mma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after_sync();
  const int warp = tid >> 5, lane = tid & 31;
  if (warp < 4) {
    for (int c0 = 0; c0 < c.N; c0 += 8) {
      float v[8];
      tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      for (int j = 0; j < 8; ++j) D[(warp * 32 + lane) * c.N + c0 + j] = v[j];
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 128);
}

int main() {
  const int KMAX = 128;
  std::vector<float> A(128 * KMAX), B(128 * KMAX);
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, 128 * 128 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * REG + 1024);
  Cfg cfgs[] = {
      {0, 0, 96, 16, 16, 1024, 16, 1024},
      {0, 1, 32, 1, 16, 1024, REG, 512},  {0, 1, 32, 1, 16, 1024, 512, REG},  {0, 1, 96, 1, 16, 1024, REG, 512},
      {1, 0, 32, 1, REG, 512, 16, 1024},  {1, 0, 32, 1, 512, REG, 16, 1024},
      {1, 1, 96, 1, REG, 512, REG, 512},  {1, 1, 96, 16, REG, 512, REG, 512},  {0, 1, 96, 16, 16, 1024, REG, 512},
      {1, 1, 96, 16, 512, REG, 512, REG},
  };
  for (auto &c : cfgs) {
    const int K = 8 * c.ksteps;
    srand(1);
    for (int i = 0; i < 128 * K; ++i) A[i] = (float)(rand() % 7 - 3);
    for (int i = 0; i < c.N * K; ++i) B[i] = (float)(rand() % 5 - 2);
    cudaMemcpy(dA, A.data(), 128 * K * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), c.N * K * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, 128 * 128 * 4);
    probe<<<1, 256, 8 * REG + 1024>>>(dA, dB, dD, c);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> D(128 * c.N);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0; int nz = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < c.N; ++n) {
        double r = 0;
        for (int k = 0; k < K; ++k) r += (double)A[m * K + k] * B[n * K + k];
        maxerr = fmax(maxerr, fabs(r - D[m * c.N + n])); maxref = fmax(maxref, fabs(r)); nz += D[m * c.N + n] != 0;
      }
    printf("a_mn=%d b_mn=%d N=%3d K=%3d lboA=%5u sboA=%5u lboB=%5u sboB=%5u : %s maxerr %.1f (maxref %.1f) nonzero %d\n", c.a_mn, c.b_mn,
           c.N, K, c.lbo_a, c.sbo_a, c.lbo_b, c.sbo_b, cudaGetErrorString(e), maxerr, maxref, nz);
    if (e != cudaSuccess) return 1;
  }
  return 0;
}
