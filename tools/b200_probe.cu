// B200 probes behind the round-2 kernels (debug aid, not part of the library):
//   1. TMA (cp.async.bulk.tensor) shared-memory layouts: SWIZZLE_NONE / SWIZZLE_128B / SWIZZLE_128B_ATOM_32B tiles are
//      dumped and compared with the offsets the UMMA descriptors of csrc/bbb_tc.cuh assume (sw128_off, mn32_off),
//      including out-of-bounds zero fill (2-D and 3-D maps).
//   2. split-K combine throughput: scalar coalesced red.global.add.f32 vs red.global.add.v4.f32 vs plain stores.
//   3. the sampling core (Philox4x32-10 -> Box-Muller -> w, mixture log p, log q) as a stand-alone loop: the
//      CUDA-core/SFU floor per weight-sample the fused kernels are held against.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/b200_probe tools/b200_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../bayesian-neural-network_b200/csrc/bbb_common.cuh"
#include "../bayesian-neural-network_b200/csrc/bbb_tc.cuh"
using namespace bbb;
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

// ---------------------------------------------------------------------------------------------------------------
// 1. TMA layouts
// ---------------------------------------------------------------------------------------------------------------
__global__ void tma_dump(const __grid_constant__ CUtensorMap map, int rank, int c0, int c1, int c2, float *out, int *status) {
  extern __shared__ uint8_t dsm[];
  __shared__ uint64_t bar;
  uint8_t *tile = (uint8_t *)(((uintptr_t)dsm + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) ((float *)tile)[i] = -777.0f;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(16384u) : "memory");
    if (rank == 2)
      asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(smem_u32(tile)), "l"(&map), "r"(smem_u32(&bar)), "r"(c0), "r"(c1) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(smem_u32(tile)), "l"(&map), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
  }
  bool ok = false;
  for (int spin = 0; spin < (1 << 22) && !ok; ++spin) ok = mbar_try_wait(smem_u32(&bar), 0);
  if (threadIdx.x == 0) *status = ok ? 1 : 0;
  __syncthreads();
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) out[i] = ((float *)tile)[i];
}

static uint32_t h_sw128_off(int row, int chunk) { return ((row >> 3) << 10) + ((row & 7) << 7) + ((chunk ^ (row & 7)) << 4); }
static uint32_t h_mn32_off(int row, int c16) { return (row << 7) + ((((c16 >> 1) ^ row) & 3) << 5) + ((c16 & 1) << 4); }

static void tma_layout_probe() {
  EncodeTiled enc = get_encode();
  const int R = 200, Cc = 100, Sx = 3;
  std::vector<float> h((size_t)Sx * R * Cc);
  for (int s = 0; s < Sx; ++s)
    for (int r = 0; r < R; ++r)
      for (int c = 0; c < Cc; ++c) h[((size_t)s * R + r) * Cc + c] = (float)(s * 100000 + r * 128 + c + 1);
  float *d, *out; int *status;
  CK(cudaMalloc(&d, h.size() * 4)); CK(cudaMalloc(&out, 16384)); CK(cudaMalloc(&status, 4));
  CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(tma_dump, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 1024));
  struct Mode { const char *name; CUtensorMapSwizzle sw; int kind; } modes[] = {
      {"SWIZZLE_NONE", CU_TENSOR_MAP_SWIZZLE_NONE, 0}, {"SWIZZLE_128B", CU_TENSOR_MAP_SWIZZLE_128B, 1},
      {"SWIZZLE_128B_ATOM_32B", CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, 2}};
  for (int rank = 2; rank <= 3; ++rank) {
    for (auto &m : modes) {
      CUtensorMap map;
      cuuint64_t dims[3] = {(cuuint64_t)Cc, (cuuint64_t)R, (cuuint64_t)Sx};
      cuuint64_t strides[2] = {(cuuint64_t)Cc * 4, (cuuint64_t)R * Cc * 4};
      cuuint32_t box[3] = {32, 128, 1}, es[3] = {1, 1, 1};
      CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, m.sw,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("tma %dd %-24s: encode failed (%d)\n", rank, m.name, (int)r); continue; }
      // box origin (col 96, row 128[, sample 1]): columns 100..127 and rows 200..255 are out of bounds -> zeros
      const int c0 = 96, c1 = 128, c2 = 1;
      tma_dump<<<1, 128, 16384 + 1024>>>(map, rank, c0, c1, c2, out, status);
      CK(cudaDeviceSynchronize());
      int st; std::vector<float> o(4096);
      CK(cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(o.data(), out, 16384, cudaMemcpyDeviceToHost));
      int bad = 0, first_bad = -1;
      for (int rr = 0; rr < 128; ++rr)
        for (int cc = 0; cc < 32; ++cc) {
          const int gr = c1 + rr, gc = c0 + cc, gs = rank == 3 ? c2 : 0;
          const float want = (gr < R && gc < Cc) ? h[((size_t)gs * R + gr) * Cc + gc] : 0.0f;
          uint32_t off = m.kind == 0 ? rr * 128 + cc * 4 : m.kind == 1 ? h_sw128_off(rr, cc >> 2) + (cc & 3) * 4 : h_mn32_off(rr, cc >> 2) + (cc & 3) * 4;
          if (o[off / 4] != want) { if (first_bad < 0) first_bad = rr * 32 + cc; ++bad; }
        }
      printf("tma %dd %-24s: completed=%d mismatches=%d of 4096 (first %d)\n", rank, m.name, st, bad, first_bad);
    }
  }
  cudaFree(d); cudaFree(out); cudaFree(status);
}

// ---------------------------------------------------------------------------------------------------------------
// 2. split-K combine: every CTA adds a [128 b][128 o] x 2-sample partial tile into y[2][128][1200]
// ---------------------------------------------------------------------------------------------------------------
template <int MODE>   // 0: scalar red, lanes along o (coalesced 128 B);  1: red.v4 along o;  2: plain scalar stores; 3: plain v4 stores
__global__ void __launch_bounds__(512) combine_kernel(float *y, int out, int n_ot) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ot = blockIdx.x % n_ot;
  const float v = 1.0f + (float)(blockIdx.x & 3);
  for (int s = 0; s < 2; ++s) {
    float *ys = y + (size_t)s * 128 * out + ot * 128;
    if (MODE == 0 || MODE == 2) {
      // warp w: lane quarter w%4 (32 consecutive o), batch columns (w/4)*32 .. +31
      const int o = (warp & 3) * 32 + lane;
      if (ot * 128 + o < out) {
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
          float *p = ys + (size_t)((warp >> 2) * 32 + j) * out + o;
          if (MODE == 0) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
          else *p = v;
        }
      }
    } else {
      // thread: row b = idx / 32, quad c = idx % 32
#pragma unroll 4
      for (int j = 0; j < 8; ++j) {
        const int idx = tid + 512 * j, b = idx >> 5, c = idx & 31;
        if (ot * 128 + c * 4 + 3 < out) {
          float *p = ys + (size_t)b * out + c * 4;
          if (MODE == 1) asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(p), "f"(v) : "memory");
          else *reinterpret_cast<float4 *>(p) = make_float4(v, v, v, v);
        }
      }
    }
  }
}

static void combine_probe() {
  const int out = 1200, n_ot = 10;
  float *y; CK(cudaMalloc(&y, (size_t)2 * 128 * out * 4));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const char *names[] = {"red.f32 coalesced", "red.v4.f32", "st.f32 coalesced", "st.v4"};
  for (int grid : {148, 296}) {
    for (int mode = 0; mode < 4; ++mode) {
      CK(cudaMemset(y, 0, (size_t)2 * 128 * out * 4));
      float best = 1e9f;
      for (int rep = 0; rep < 6; ++rep) {
        CK(cudaEventRecord(e0));
        for (int it = 0; it < 10; ++it) {
          if (mode == 0) combine_kernel<0><<<grid, 512>>>(y, out, n_ot);
          if (mode == 1) combine_kernel<1><<<grid, 512>>>(y, out, n_ot);
          if (mode == 2) combine_kernel<2><<<grid, 512>>>(y, out, n_ot);
          if (mode == 3) combine_kernel<3><<<grid, 512>>>(y, out, n_ot);
        }
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
      }
      printf("combine grid=%3d %-18s: %.2f us per launch (%.1f M floats per launch)\n", grid, names[mode], best * 100.0f,
             grid * 32768 / 1e6);
    }
  }
  cudaFree(y);
}

// ---------------------------------------------------------------------------------------------------------------
// 3. the sampling core
// ---------------------------------------------------------------------------------------------------------------
struct CoreArgs {
  const float *mu, *rho;
  float *sink;
  int n_quads, iters;
  uint32_t rk0[10], rk1[10];
  PriorDev prior;
};

__device__ __forceinline__ uint4 philox_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const CoreArgs &a) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ a.rk0[r], n2 = (uint32_t)(p0 >> 32) ^ c3 ^ a.rk1[r];
    c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
  }
  return make_uint4(c0, c1, c2, c3);
}
// low 23 bits -> float in [1, 2): (x & 0x7fffff) | one, ONE lop3 (the 1.0f pattern comes in a register so that the
// compiler cannot split the expression into two immediates)
__device__ __forceinline__ float bits_to_12(uint32_t x, uint32_t one) {
  uint32_t r;
  asm("lop3.b32 %0, %1, 0x007fffff, %2, 0xEA;" : "=r"(r) : "r"(x), "r"(one));
  return __uint_as_float(r);
}
__device__ __forceinline__ void bm_lean(uint32_t a, uint32_t b, uint32_t one, float &z0, float &z1) {
  const float u = bits_to_12(a, one), v = bits_to_12(b, one);
  const float t = -1.3862943611198906f * lg2_approx(2.0f - u);        // -2 ln(u'), u' in (0, 1]
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
  float s, c;
  __sincosf(v * 6.283185307179586f, &s, &c);                          // sin / cos are 2 pi periodic: v in [1, 2) is fine
  z0 = r * s; z1 = r * c;
}
// round-to-nearest TF32: the tensor core truncates the low 13 bits itself, so adding half a TF32 ulp is enough
__device__ __forceinline__ float tf32_round(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }
// softplus without the rho > 15 branch (e^rho overflows only where the reference's own log1p(exp(rho)) does)
__device__ __forceinline__ float softplus_lean(float rho) {
  const float t = ex2_approx(rho * 1.4426950408889634f);
  float ser = fmaf(t, -0.16666667f, 0.2f);
  ser = fmaf(t, ser, -0.25f);
  ser = fmaf(t, ser, 0.33333334f);
  ser = fmaf(t, ser, -0.5f);
  ser = fmaf(t, ser, 1.0f);
  const float lg = 0.6931471805599453f * lg2_approx(1.0f + t);
  return t < 0.125f ? ser * t : lg;
}

template <int VARIANT>   // 0: current library functions (bbb_common.cuh);  1: lean variant
__global__ void __launch_bounds__(256) core_kernel(const CoreArgs a) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  float lp[2] = {0.f, 0.f}, lq[2] = {0.f, 0.f}, ws = 0.f;
  RngDev rng{}; rng.key0 = 17; rng.key1 = 3; rng.step = 5;
  uint32_t one;
  asm volatile("mov.b32 %0, 0x3f800000;" : "=r"(one));
  for (int it = 0; it < a.iters; ++it) {
    const int q = (tid + it * nth) % a.n_quads;
    const float4 m = __ldg(reinterpret_cast<const float4 *>(a.mu) + q), r4 = __ldg(reinterpret_cast<const float4 *>(a.rho) + q);
    const float mu[4] = {m.x, m.y, m.z, m.w}, rho[4] = {r4.x, r4.y, r4.z, r4.w};
    float sg[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) sg[c] = VARIANT == 0 ? softplus_fast(rho[c]) : softplus_lean(rho[c]);
    const float lsg = logsigma_quad_fast(sg);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      float ep[4], w[4];
      if (VARIANT == 0) {
        philox_normal4(rng, 2, (uint32_t)s, (uint32_t)q, ep);
      } else {
        const uint4 rr = philox_rk((uint32_t)q, (uint32_t)s, 2u, 5u, a);
        bm_lean(rr.x, rr.y, one, ep[0], ep[1]);
        bm_lean(rr.z, rr.w, one, ep[2], ep[3]);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) w[c] = fmaf(sg[c], ep[c], mu[c]);
      if (VARIANT == 0) ws += to_tf32(w[0]) + to_tf32(w[1]) + to_tf32(w[2]) + to_tf32(w[3]);
      else ws += tf32_round(w[0]) + tf32_round(w[1]) + tf32_round(w[2]) + tf32_round(w[3]);
      lp[s] += logp_quad_fast(a.prior, w);
      lq[s] += -4.0f * kHalfLog2Pi - lsg - 0.5f * (ep[0] * ep[0] + ep[1] * ep[1] + ep[2] * ep[2] + ep[3] * ep[3]);
    }
  }
  a.sink[tid] = lp[0] + lp[1] + lq[0] + lq[1] + ws;
}

static void core_probe() {
  const int n_quads = 360000;   // 1200 x 1200 / 4
  std::vector<float> hm(n_quads * 4), hr(n_quads * 4);
  srand(3);
  for (auto &v : hm) v = -0.2f + 0.4f * (rand() / (float)RAND_MAX);
  for (auto &v : hr) v = -5.0f + (rand() / (float)RAND_MAX);
  CoreArgs a{};
  float *mu, *rho, *sink;
  CK(cudaMalloc(&mu, hm.size() * 4)); CK(cudaMalloc(&rho, hr.size() * 4)); CK(cudaMalloc(&sink, 148 * 16 * 256 * 4));
  CK(cudaMemcpy(mu, hm.data(), hm.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(rho, hr.data(), hr.size() * 4, cudaMemcpyHostToDevice));
  a.mu = mu; a.rho = rho; a.sink = sink; a.n_quads = n_quads;
  uint32_t k0 = 17, k1 = 3;
  for (int r = 0; r < 10; ++r) { a.rk0[r] = k0; a.rk1[r] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
  bbb_prior pr{BBB_PRIOR_MIXTURE, 0.5f, 1.0f, 3.3546e-4f};
  a.prior = make_prior_dev(&pr);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int variant = 0; variant < 2; ++variant) {
    for (int cps : {2, 4, 8}) {                     // resident CTAs of 256 threads per SM
      const int grid = 148 * cps;
      a.iters = (n_quads + grid * 256 - 1) / (grid * 256) * 8;      // 8 passes over a 1200 x 1200 layer
      float best = 1e9f;
      for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        if (variant == 0) core_kernel<0><<<grid, 256>>>(a); else core_kernel<1><<<grid, 256>>>(a);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
      }
      const double ws = (double)a.iters * grid * 256 * 4 * 2;   // weight-samples processed
      printf("core variant=%d ctas/sm=%d: %.1f us for %.2f M weight-samples -> %.3f ns per 1000 -> a 1200x1200 layer, S=2: %.2f us\n",
             variant, cps, best * 1000.0, ws / 1e6, best * 1e6 / ws * 1000.0, best * 1000.0 * 2.88e6 / ws);
    }
  }
  CK(cudaDeviceSynchronize());
}

int main(int argc, char **argv) {
  CK(cudaSetDevice(0));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s, %d SMs, L2 %d MB\n", p.name, p.multiProcessorCount, p.l2CacheSize >> 20);
  tma_layout_probe();
  combine_probe();
  core_probe();
  return 0;
}
