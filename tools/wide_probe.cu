// B200 probes behind the wide-layer (batch-resident) kernels (debug aid, not part of the library):
//   1. L2 -> shared-memory streaming rate of TMA tile loads when every SM streams (the forward's activation tiles:
//      [512 batch rows][32 k] fp32 = 64 KB per stage out of a 64 MB L2-resident matrix), unicast and with the tile
//      loaded once per CTA pair and multicast to both;
//   2. the cluster protocol the kernels share sampled weight tiles with: bulk shared::cta -> shared::cluster copies
//      completing on the destination CTA's mbarrier, and tcgen05.commit multicast to every CTA of the cluster.
// Every wait is bounded: a protocol mistake reports "timeout" instead of hanging the GPU.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/wide_probe tools/wide_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../bayesian-neural-network_b200/csrc/bbb_tc.cuh"
using namespace bbb::tc;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiled get_encode() {
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &q));
  if (q != cudaDriverEntryPointSuccess || !fn) { printf("cuTensorMapEncodeTiled not found\n"); exit(1); }
  return (EncodeTiled)fn;
}

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void remote_arrive(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool bounded_wait(uint32_t bar, uint32_t parity, int spins = 1 << 24) {
  for (int i = 0; i < spins; ++i)
    if (mbar_try_wait(bar, parity)) return true;
  return false;
}

// ---------------------------------------------------------------------------------------------------------------
// 1. streaming rate
// ---------------------------------------------------------------------------------------------------------------
constexpr int NST = 3, TILE = 65536;
template <int MC>   // 0: every CTA loads its own tile; 1: CTA pairs share a tile (each loads half, multicast to both)
__global__ void __launch_bounds__(128, 1) stream_kernel(const __grid_constant__ CUtensorMap map, int nkb, int reps, int *status) {
  extern __shared__ uint8_t dsm[];
  __shared__ uint64_t full[NST], empty[NST];
  uint8_t *tiles = (uint8_t *)(((uintptr_t)dsm + 1023) & ~(uintptr_t)1023);
  const uint32_t rank = MC ? cluster_rank() : 0;
  const int tile_id = MC ? blockIdx.x / 2 : blockIdx.x;
  const int b0 = (tile_id % 8) * 512;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), MC ? 2 : 1); }
    mbar_fence_init();
  }
  __syncthreads();
  if (MC) cluster_sync_all();
  const int total = nkb * reps;
  if (threadIdx.x == 0) {
    for (int it = 0; it < total; ++it) {
      const int st = it % NST;
      if (it >= NST && !bounded_wait(smem_u32(&empty[st]), ((it / NST) - 1) & 1)) { *status = 2; break; }
      const uint32_t dst = smem_u32(tiles + st * TILE), bar = smem_u32(&full[st]);
      const int k0 = (it % nkb) * 32;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)TILE) : "memory");
      if (!MC) {
        for (int h = 0; h < 2; ++h)
          asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(dst + h * 32768), "l"(&map), "r"(bar), "r"(k0), "r"(b0 + 256 * h) : "memory");
      } else {
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                     ::"r"(dst + rank * 32768), "l"(&map), "r"(bar), "r"(k0), "r"(b0 + 256 * (int)rank), "h"((uint16_t)3) : "memory");
      }
    }
  } else if (threadIdx.x == 32) {
    for (int it = 0; it < total; ++it) {
      const int st = it % NST;
      if (!bounded_wait(smem_u32(&full[st]), (it / NST) & 1)) { *status = 3; break; }
      if (!MC) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[st])) : "memory");
      else { remote_arrive(mapa(smem_u32(&empty[st]), 0)); remote_arrive(mapa(smem_u32(&empty[st]), 1)); }
    }
  }
  __syncthreads();
  if (MC) cluster_sync_all();
}

static void stream_probe() {
  EncodeTiled enc = get_encode();
  const int64_t R = 4096, Cc = 4096;
  float *d; int *status;
  CK(cudaMalloc(&d, R * Cc * 4)); CK(cudaMemset(d, 0, R * Cc * 4)); CK(cudaMalloc(&status, 4));
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)Cc, (cuuint64_t)R}, strides[1] = {(cuuint64_t)Cc * 4};
  cuuint32_t box[2] = {32, 256}, es[2] = {1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return; }
  const int dyn = NST * TILE + 1024, nkb = 128, reps = 4;
  CK(cudaFuncSetAttribute(stream_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
  CK(cudaFuncSetAttribute(stream_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int mc = 0; mc < 2; ++mc) {
    for (int grid : {148, 128, 64}) {
      float best = 1e30f;
      for (int rep = 0; rep < 4; ++rep) {
        CK(cudaMemset(status, 0, 4));
        CK(cudaEventRecord(e0));
        if (!mc) stream_kernel<0><<<grid, 128, dyn>>>(map, nkb, reps, status);
        else {
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = dyn;
          cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim = {2, 1, 1};
          cfg.attrs = at; cfg.numAttrs = 1;
          CK(cudaLaunchKernelEx(&cfg, stream_kernel<1>, map, nkb, reps, status));
        }
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
      }
      int st; CK(cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost));
      const double bytes_sm = (double)nkb * reps * TILE;          // landed in every SM
      const double tbs = bytes_sm * grid / (best * 1e-3) / 1e12;
      printf("stream %s grid=%3d: %.3f ms, %.2f TB/s into shared memory in total, %.1f B/clk per SM (1.92 GHz)%s\n",
             mc ? "pair-multicast" : "unicast       ", grid, best, tbs, bytes_sm / (best * 1e-3) / 1.92e9, st ? "  TIMEOUT" : "");
    }
  }
  cudaFree(d); cudaFree(status);
}

// ---------------------------------------------------------------------------------------------------------------
// 2. cluster protocol: bulk DSMEM copies + multicast commit
// ---------------------------------------------------------------------------------------------------------------
constexpr int CS = 4, PIECE = 4096;
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(128, 1) proto_kernel(int *result) {
  __shared__ __align__(1024) uint8_t tile[CS * PIECE];     // 128 rows x 128 B, one 32-row piece per CTA
  __shared__ __align__(1024) uint8_t btile[8192];
  __shared__ uint64_t w_full, w_empty, acc;
  __shared__ uint32_t tbase;
  const uint32_t rank = cluster_rank();
  const int tid = threadIdx.x;
  if (tid < 32) tmem_alloc(smem_u32(&tbase), 64);
  if (tid == 32) {
    mbar_init(smem_u32(&w_full), 1); mbar_init(smem_u32(&w_empty), CS); mbar_init(smem_u32(&acc), 1);
    mbar_fence_init();
  }
  for (int i = tid; i < CS * PIECE / 4; i += 128) ((float *)tile)[i] = -1.0f;
  for (int i = tid; i < 2048; i += 128) ((float *)btile)[i] = 0.0f;
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  cluster_sync_all();
  int ok = 1;
  for (int round = 0; round < 3 && ok; ++round) {
    // my piece: rows 32 rank .. +31, value = 1000 round + rank (as a K-major operand tile it is just bytes here)
    for (int i = tid; i < PIECE / 4; i += 128) ((float *)(tile + rank * PIECE))[i] = (float)(1000 * round + rank);
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&w_full)), "r"((uint32_t)((CS - 1) * PIECE)) : "memory");
      for (uint32_t p = 0; p < CS; ++p) {
        if (p == rank) continue;
        const uint32_t dst = mapa(smem_u32(tile + rank * PIECE), p), bar = mapa(smem_u32(&w_full), p);
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "r"(smem_u32(tile + rank * PIECE)), "r"((uint32_t)PIECE), "r"(bar) : "memory");
      }
    }
    if (!bounded_wait(smem_u32(&w_full), round & 1)) { ok = 0; if (tid == 0) result[8 + rank] = 100 + round; }
    __syncthreads();
    if (ok) {
      for (int i = tid; i < CS * PIECE / 4; i += 128) {
        const float want = (float)(1000 * round + i / (PIECE / 4));
        if (((float *)tile)[i] != want) { ok = 0; result[8 + rank] = 200 + round; }
      }
    }
    ok = __syncthreads_and(ok);
    // one MMA reading the tile, then a commit multicast to the w_empty barrier of every CTA of the cluster
    if (tid == 0) {
      tc_fence_after_sync();
      mma_tf32(tbase, smem_desc_sw128(smem_u32(tile)), smem_desc_sw128(smem_u32(btile)), idesc_tf32(128, 64), 0u);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                   ::"r"(smem_u32(&w_empty)), "h"((uint16_t)((1 << CS) - 1)) : "memory");
    }
    if (!bounded_wait(smem_u32(&w_empty), round & 1)) { ok = 0; if (tid == 0) result[8 + rank] = 300 + round; }
    ok = __syncthreads_and(ok);
  }
  if (tid == 0) result[rank] = ok;
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (tid < 32) tmem_dealloc(tbase, 64);
}

static void proto_probe() {
  int *res; CK(cudaMalloc(&res, 64)); CK(cudaMemset(res, 0, 64));
  proto_kernel<<<CS, 128>>>(res);
  cudaError_t e = cudaDeviceSynchronize();
  int h[16];
  if (e != cudaSuccess) { printf("proto: CUDA error %s\n", cudaGetErrorString(e)); return; }
  CK(cudaMemcpy(h, res, 64, cudaMemcpyDeviceToHost));
  printf("proto (cluster of %d: bulk DSMEM copies + multicast commit): ok = %d %d %d %d, codes = %d %d %d %d\n", CS, h[0], h[1], h[2], h[3],
         h[8], h[9], h[10], h[11]);
  cudaFree(res);
}

int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s, %d SMs\n", p.name, p.multiProcessorCount);
  proto_probe();
  stream_probe();
  return 0;
}
