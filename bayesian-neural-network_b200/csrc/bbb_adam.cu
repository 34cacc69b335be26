// Multi-tensor Adam: torch.optim.Adam.step() for every parameter tensor of the network in ONE launch
// (SURVEY 8f-1).  The stock optimiser issues several foreach kernels per step; at the MNIST-shape config
// those passes cost more than the whole fused forward+backward.  HBM-bound: 16 B read + 12 B written per
// parameter (p, g, m, v in; p, m, v out), 16-byte vector accesses, grid = a multiple of the 148 SMs.
// The update rule is exactly torch's (non-amsgrad, no weight decay):
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
#include <string.h>

#include "bbb_common.cuh"
#include "bbb_mlp.h"

namespace bbb {
namespace {

constexpr int kMaxTensors = 32;
struct AdamArgs {
  float *p[kMaxTensors];
  const float *g[kMaxTensors];
  float *m[kMaxTensors];
  float *v[kMaxTensors];
  int64_t size[kMaxTensors];  // elements
  int64_t qend[kMaxTensors];  // cumulative quads (ceil(size/4)) up to and including this tensor
  bool vec[kMaxTensors];      // all four pointers 16-byte aligned
  int n;
  double lr, b1, b2;  // torch forms 1-b, the bias corrections and the step size in double: so does the kernel
  float eps;
  uint32_t step;
  const uint32_t *step_dev;
  const float *lr_scale_dev;
};

// One 16-byte quad of one tensor: where it lives and whether it can be moved with vector accesses.
struct QuadRef {
  float *p, *m, *v;
  const float *g;
  int valid;     // elements of the quad inside the tensor (4 unless it is the tensor's tail)
  bool vec;
};
__device__ __forceinline__ QuadRef locate(const AdamArgs &a, int64_t q, int &ti) {
  while (q >= a.qend[ti]) ++ti;
  const int64_t e = (q - (ti ? a.qend[ti - 1] : 0)) * 4;
  QuadRef r;
  r.p = a.p[ti] + e; r.m = a.m[ti] + e; r.v = a.v[ti] + e; r.g = a.g[ti] + e;
  r.valid = (int)min((int64_t)4, a.size[ti] - e);
  r.vec = r.valid == 4 && a.vec[ti];
  return r;
}
// the SFU's sqrt and reciprocal are accurate to 1 ulp-level (2^-23 relative); the update term they feed is lr times
// smaller than the parameter, so the result equals torch's to far below fp32 round-off of p (tests: rtol 2e-6)
__device__ __forceinline__ void adam4(float4 &P, const float4 &G, float4 &M, float4 &V, const AdamConst &c, float ibc) {
  adam1_fast(P.x, G.x, M.x, V.x, c, ibc);
  adam1_fast(P.y, G.y, M.y, V.y, c, ibc);
  adam1_fast(P.z, G.z, M.z, V.z, c, ibc);
  adam1_fast(P.w, G.w, M.w, V.w, c, ibc);
}

__global__ void __launch_bounds__(256) adam_kernel(const AdamArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ AdamConst cs;
  if (threadIdx.x == 0) cs = adam_consts(a.lr, a.b1, a.b2, a.eps, a.step, a.step_dev, a.lr_scale_dev);
  __syncthreads();
  const AdamConst c = cs;
  const float ibc = 1.0f / c.bc2_sqrt;
  const int64_t total = a.qend[a.n - 1], stride = (int64_t)gridDim.x * blockDim.x;
  int t0 = 0, t1 = 0;
  // two quads per thread and iteration: all eight 16-byte loads are in flight before the first update is formed
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += 2 * stride) {
    const QuadRef r0 = locate(a, q, t0);
    const bool two = q + stride < total;
    QuadRef r1 = r0;
    if (two) r1 = locate(a, q + stride, t1);
    if (r0.vec && r1.vec) {
      float4 P0 = *reinterpret_cast<float4 *>(r0.p), M0 = *reinterpret_cast<float4 *>(r0.m), V0 = *reinterpret_cast<float4 *>(r0.v);
      const float4 G0 = *reinterpret_cast<const float4 *>(r0.g);
      float4 P1 = *reinterpret_cast<float4 *>(r1.p), M1 = *reinterpret_cast<float4 *>(r1.m), V1 = *reinterpret_cast<float4 *>(r1.v);
      const float4 G1 = *reinterpret_cast<const float4 *>(r1.g);
      adam4(P0, G0, M0, V0, c, ibc);
      *reinterpret_cast<float4 *>(r0.p) = P0;
      *reinterpret_cast<float4 *>(r0.m) = M0;
      *reinterpret_cast<float4 *>(r0.v) = V0;
      if (two) {
        adam4(P1, G1, M1, V1, c, ibc);
        *reinterpret_cast<float4 *>(r1.p) = P1;
        *reinterpret_cast<float4 *>(r1.m) = M1;
        *reinterpret_cast<float4 *>(r1.v) = V1;
      }
    } else {
      for (int k = 0; k < (two ? 2 : 1); ++k) {
        const QuadRef &r = k ? r1 : r0;
        for (int j = 0; j < r.valid; ++j) {
          float P = r.p[j], M = r.m[j], V = r.v[j];
          adam1_fast(P, r.g[j], M, V, c, ibc);
          r.p[j] = P; r.m[j] = M; r.v[j] = V;
        }
      }
    }
  }
}

// ==================================================================================================
// Multi-GPU step: gradient reduce-scatter + Adam + parameter all-gather in ONE kernel over NVLink peer memory.
// Rank r owns the slice [r n/W, (r+1) n/W) of the flat parameter space.  For its slice it sums the W ranks' gradient
// buckets (peer loads, fixed rank order), applies the Adam update with its local m / v, and stores the new
// parameters into every rank's parameter buffer (peer stores).  Two flag barriers in peer memory bracket it:
//   entry: every rank's gradients are complete (its backward has finished) before anyone reads them;
//   exit : every rank has finished reading my gradients and writing my parameters before my kernel completes,
//          so the next step on this stream may overwrite the bucket and read the parameters.
// Flags are monotonically increasing call counts (never reset).  No NCCL, no extra launch, no extra pass over the
// gradients; each rank touches 1/W of the optimiser state.
// ==================================================================================================
struct PeerArgs {
  int world, rank;
  const float *g[BBB_MAX_PEERS];
  float *p[BBB_MAX_PEERS];
  uint32_t *flags[BBB_MAX_PEERS];
  uint32_t *epoch, *done;
  float *m, *v;
  int64_t n;
  double lr, b1, b2;
  float eps;
  uint32_t step;
  const uint32_t *step_dev;
  const float *lr_scale_dev;
  unsigned long long *tl;      // debug (bbb_debug_set_timeline): phase stamps of block 0 / the last block, or NULL
  const float *mc_g;           // NVLS multicast mappings of the gradient / parameter buffers, or NULL
  float *mc_p;
};

// sum over every rank's copy of 16 bytes, formed by the NVSwitch; store of 16 bytes to every rank's copy
__device__ __forceinline__ float4 multimem_ld_sum4(const float *mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st4(float *mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

__device__ __forceinline__ void pstamp(unsigned long long *tl, int slot) {
  if (tl) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    tl[slot] = t;
  }
}
__device__ __forceinline__ void st_flag_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_flag_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Bounded wait on a peer's flag: a rank that skips a step or dies must not leave the other GPUs spinning for ever inside a
// kernel.  After ~2^24 polls (seconds) the waiter gives up, raises the error word (done[1], checked by the host:
// PeerShardedAdam.check_health) and lets the kernel finish with whatever it has.
__device__ __forceinline__ void wait_flag(const uint32_t *f, uint32_t e, uint32_t *err) {
  for (uint32_t spin = 0; (int32_t)(ld_flag_sys(f) - e) < 0; ++spin) {
    if (spin > (1u << 24)) { atomicExch(err, 1u); break; }
    __nanosleep(40);
  }
}

__global__ void __launch_bounds__(256) peer_adam_kernel(const PeerArgs a) {
  pdl_launch_dependents();
  pdl_wait();                         // this rank's backward (the previous kernels of the stream) is complete
  __shared__ AdamConst cs;
  __shared__ uint32_t last_s;
  const int W = a.world, tid = threadIdx.x;
  unsigned long long *tl0 = (a.tl && blockIdx.x == 0 && tid == 0) ? a.tl : nullptr;
  pstamp(tl0, 0);
  const uint32_t e = *reinterpret_cast<volatile uint32_t *>(a.epoch) + 1u;   // this call's number (same in every block)
  if (tid == 0) cs = adam_consts(a.lr, a.b1, a.b2, a.eps, a.step, a.step_dev, a.lr_scale_dev);
  // ---- entry barrier: block 0 announces this rank, every block waits for all ranks
  if (blockIdx.x == 0 && tid < W) {
    __threadfence_system();
    st_flag_sys(a.flags[tid] + a.rank, e);
  }
  if (tid < W) {
    wait_flag(a.flags[a.rank] + tid, e, a.done + 1);
  }
  __syncthreads();
  pstamp(tl0, 1);                       // every rank's gradients are complete
  const AdamConst c = cs;
  const float inv_w = 1.0f / (float)W;
  // ---- this rank's slice, in 16-byte quads; the last rank also takes the n % 4 tail
  const int64_t nq = a.n >> 2;
  const int64_t q_lo = nq * a.rank / W, q_hi = nq * (a.rank + 1) / W;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // (Measured with tools/time_peer_adam.py at 2 GPUs: entry barrier 6 us, reads + update 19-25 us, drain of the peer
  //  stores 12-14 us, exit barrier 4-6 us.  A two-deep software pipeline that overlaps the inbound reads with the outbound
  //  writes took the same 36 us for the middle two: not kept.)
  const bool nvls_ld = a.mc_g != nullptr, nvls_st = a.mc_p != nullptr;
  for (int64_t q = q_lo + (int64_t)blockIdx.x * blockDim.x + tid; q < q_hi; q += stride) {
    float4 G = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nvls_ld) {
      G = multimem_ld_sum4(a.mc_g + 4 * q);
    } else {
#pragma unroll
      for (int k = 0; k < BBB_MAX_PEERS; ++k) {
        if (k < W) {
          const float4 t = __ldcg(reinterpret_cast<const float4 *>(a.g[k]) + q);   // peer memory: never through L1
          G.x += t.x; G.y += t.y; G.z += t.z; G.w += t.w;
        }
      }
    }
    G.x *= inv_w; G.y *= inv_w; G.z *= inv_w; G.w *= inv_w;
    float4 P = reinterpret_cast<const float4 *>(a.p[a.rank])[q];
    float4 M = reinterpret_cast<const float4 *>(a.m)[q], V = reinterpret_cast<const float4 *>(a.v)[q];
    adam1(P.x, G.x, M.x, V.x, c);
    adam1(P.y, G.y, M.y, V.y, c);
    adam1(P.z, G.z, M.z, V.z, c);
    adam1(P.w, G.w, M.w, V.w, c);
    reinterpret_cast<float4 *>(a.m)[q] = M;
    reinterpret_cast<float4 *>(a.v)[q] = V;
    if (nvls_st) {
      multimem_st4(a.mc_p + 4 * q, P);
    } else {
#pragma unroll
      for (int k = 0; k < BBB_MAX_PEERS; ++k)
        if (k < W) reinterpret_cast<float4 *>(a.p[k])[q] = P;
    }
  }
  if (a.rank == W - 1 && blockIdx.x == 0 && tid < (a.n & 3)) {
    const int64_t i = (nq << 2) + tid;
    float G = 0.0f;
    for (int k = 0; k < W; ++k) G += __ldcg(a.g[k] + i);
    G *= inv_w;
    float P = a.p[a.rank][i], M = a.m[i], V = a.v[i];
    adam1(P, G, M, V, c);
    a.m[i] = M; a.v[i] = V;
    for (int k = 0; k < W; ++k) a.p[k][i] = P;
  }
  // ---- exit barrier: the last block to finish announces this rank and waits for the others
  pstamp(tl0, 2);                       // block 0: loads, update and stores issued
  __threadfence_system();             // this thread's peer stores are performed
  __syncthreads();
  pstamp(tl0, 3);                       // block 0: its peer stores are performed
  if (tid == 0) last_s = atomicAdd(a.done, 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (last_s) {
    unsigned long long *tl1 = (a.tl && tid == 0) ? a.tl : nullptr;
    pstamp(tl1, 4);                     // the last block of this rank has finished
    if (tid < W) {
      __threadfence_system();
      st_flag_sys(a.flags[tid] + W + a.rank, e);
      wait_flag(a.flags[a.rank] + W + tid, e, a.done + 1);
    }
    __syncthreads();
    pstamp(tl1, 5);                     // every rank has finished
    if (tid == 0) { *a.done = 0u; *a.epoch = e; }
  }
}

}  // namespace
}  // namespace bbb

using namespace bbb;

extern "C" int bbb_enable_peer_access(int32_t peer_device) {
  int cur = -1;
  BBB_CHECK_CUDA(cudaGetDevice(&cur));
  if (cur == peer_device) return BBB_OK;
  int can = 0;
  BBB_CHECK_CUDA(cudaDeviceCanAccessPeer(&can, cur, peer_device));
  if (!can) return fail(BBB_EUNSUPPORTED, "bbb_enable_peer_access: device %d cannot access device %d", cur, peer_device);
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return BBB_OK; }
  BBB_CHECK_CUDA(e);
  return BBB_OK;
}

extern "C" int bbb_ipc_open(const void *handle, int64_t offset_bytes, void **out_ptr) {
  BBB_CHECK_ARG(handle && out_ptr && offset_bytes >= 0, "null pointer or negative offset");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void *base = nullptr;
  // opened with the CALLER's device current: the mapping is then usable by that device's kernels (peer access to
  // the exporting device is enabled on demand)
  BBB_CHECK_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  *out_ptr = static_cast<char *>(base) + offset_bytes;
  return BBB_OK;
}

extern "C" int bbb_adam_step_peer(const bbb_peer_comm *comm, float *exp_avg, float *exp_avg_sq, int64_t n, double lr,
                                  double beta1, double beta2, double eps, uint32_t step, const uint32_t *step_dev,
                                  const float *lr_scale_dev, void *stream) {
  BBB_CHECK_ARG(comm && exp_avg && exp_avg_sq, "null pointer");
  BBB_CHECK_ARG(comm->world >= 1 && comm->world <= BBB_MAX_PEERS && comm->rank >= 0 && comm->rank < comm->world,
                "1 <= world <= 8, 0 <= rank < world");
  BBB_CHECK_ARG(comm->epoch && comm->done_blocks, "null epoch / counter word");
  BBB_CHECK_ARG(n >= 0, "negative size");
  BBB_CHECK_ARG(step + (step_dev ? 1u : 0u) >= 1u, "Adam step is 1-based");
  PeerArgs a{};
  a.world = comm->world; a.rank = comm->rank;
  for (int k = 0; k < comm->world; ++k) {
    BBB_CHECK_ARG(comm->grads[k] && comm->params[k] && comm->flags[k], "null peer pointer");
    BBB_CHECK_ARG(((reinterpret_cast<uintptr_t>(comm->grads[k]) | reinterpret_cast<uintptr_t>(comm->params[k])) & 15u) == 0,
                  "peer buffers must be 16-byte aligned");
    a.g[k] = comm->grads[k]; a.p[k] = comm->params[k]; a.flags[k] = comm->flags[k];
  }
  BBB_CHECK_ARG(((reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15u) == 0,
                "optimiser state must be 16-byte aligned");
  a.epoch = comm->epoch; a.done = comm->done_blocks;
  a.m = exp_avg; a.v = exp_avg_sq; a.n = n;
  a.lr = lr; a.b1 = beta1; a.b2 = beta2; a.eps = (float)eps; a.step = step; a.step_dev = step_dev;
  a.lr_scale_dev = lr_scale_dev;
  a.tl = peer_timeline();
  a.mc_g = comm->mc_grads; a.mc_p = comm->mc_params;
  BBB_CHECK_ARG(((reinterpret_cast<uintptr_t>(a.mc_g) | reinterpret_cast<uintptr_t>(a.mc_p)) & 15u) == 0, "multicast mappings must be 16-byte aligned");
  // every block must be able to run while block 0 is still waiting at the entry barrier, and the grid of every rank
  // must make progress independently: at most 4 blocks per SM, so the grid is always co-resident
  const int64_t slice_q = (n >> 2) / comm->world + 1;
  int64_t blocks = (slice_q + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  BBB_CHECK_CUDA(launch_pdl(peer_adam_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, a));
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}

extern "C" int bbb_adam_step(int32_t n_tensors, float *const *params, const float *const *grads, float *const *exp_avg,
                             float *const *exp_avg_sq, const int64_t *sizes, double lr, double beta1, double beta2,
                             double eps, uint32_t step, const uint32_t *step_dev, const float *lr_scale_dev,
                             void *stream) {
  BBB_CHECK_ARG(n_tensors >= 0 && n_tensors <= kMaxTensors, "0 <= n_tensors <= 32");
  BBB_CHECK_ARG(n_tensors == 0 || (params && grads && exp_avg && exp_avg_sq && sizes), "null table");
  BBB_CHECK_ARG(step + (step_dev ? 1u : 0u) >= 1u, "Adam step is 1-based");
  AdamArgs a{};
  int64_t q = 0;
  int n = 0;
  for (int i = 0; i < n_tensors; ++i) {
    BBB_CHECK_ARG(sizes[i] >= 0, "negative size");
    if (sizes[i] == 0) continue;
    BBB_CHECK_ARG(params[i] && grads[i] && exp_avg[i] && exp_avg_sq[i], "null tensor pointer");
    a.p[n] = params[i]; a.g[n] = grads[i]; a.m[n] = exp_avg[i]; a.v[n] = exp_avg_sq[i];
    a.size[n] = sizes[i];
    q += (sizes[i] + 3) / 4;
    a.qend[n] = q;
    a.vec[n] = ((reinterpret_cast<uintptr_t>(params[i]) | reinterpret_cast<uintptr_t>(grads[i]) |
                 reinterpret_cast<uintptr_t>(exp_avg[i]) | reinterpret_cast<uintptr_t>(exp_avg_sq[i])) & 15u) == 0;
    ++n;
  }
  if (n == 0) return BBB_OK;
  a.n = n; a.lr = lr; a.b1 = beta1; a.b2 = beta2; a.eps = (float)eps; a.step = step; a.step_dev = step_dev;
  a.lr_scale_dev = lr_scale_dev;
  int64_t blocks = (q + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 4;   // one resident wave (4 CTAs of 256 threads x 56 registers per SM)
  if (blocks > cap) blocks = cap;
  BBB_CHECK_CUDA(launch_pdl(adam_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, a));
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}
