// The HEAD of the network: a weight-sampling Bayesian linear layer with a narrow output (out <= 16; MNIST-shape
// 1200 -> 10, regression / bandit hidden -> 1) together with everything that hangs off its output.  Such a layer has
// a few thousand weights: neither HBM nor the tensor pipe matters, only the length of the chain of dependent phases
// on the step's critical path.  Two kernels, both exact fp32 (FMA contractions, precise log-densities), both
// deterministic (no atomics on tensors):
//
//   head_fwd   = BayesianLinear.forward (networks.py:73-88) for all S samples
//                + get_nll (networks.py:183-190) and its gradient w.r.t. the outputs
//                + the ELBO assembly (networks.py:205-209)                              -- ONE launch instead of 3.
//     Thread-block clusters of 16 CTAs (8 where the device cannot co-schedule 16) of 512 threads.  A cluster owns up
//     to 2 samples x a power-of-two block of batch rows (MNIST shape: both samples x 128 rows = ONE cluster, so
//     nothing is synchronised through global memory).  Its CTAs split K: CTA c samples W_s[:, k range c] into shared
//     memory -- every weight quad is formed exactly once, one per thread -- while its activation quads are already
//     in flight, forms the partial dot products of its k range, and the partials are summed over the cluster
//     through distributed shared memory in a fixed order: CTA c finishes rows [c R/16, (c+1) R/16) with 16 lanes per
//     row (one per output column): bias, y, the likelihood term and d nll / d y.  The CTA sums of (log p, log q, nll)
//     go to rank 0 through distributed shared memory; rank 0 adds them to the fp64 accumulators and, when it is the
//     only (or the last) cluster, assembles the four ELBO scalars.
//
//   head_bwd   = the layer's backward: grid over 4-aligned ranges of <= 16 input columns.  A CTA owns W[:, i_lo:i_hi)
//     completely.  A pass covers up to 256 (sample, batch-row) pairs, ONE PER THREAD (all samples at once when
//     S B <= 256): the thread keeps its dz row and its x columns in registers, so dX_s[b][i] = sum_o dz_s[b][o] W_s[o][i]
//     is a register contraction against the broadcast W tile (plain 16-byte stores, optional (x > 0) mask), while
//     G_s[o][i] = sum_b dz_s[b][o] x_s[b][i] is formed from shared memory by 4 batch groups x 64 (sample, o, quad)
//     items.  The item threads sample their weight quad once (while the loads are in flight), finish it with the
//     analytic mu/rho-gradient epilogue, and the quad's owner adds up the samples.
#include "bbb_common.cuh"
#include "bbb_kernels.h"

namespace bbb {
namespace {

constexpr int NO = 16;     // widest output handled

struct HQuad {
  float mu[4], sg[4], ep[4], w[4];
};
// Parameters and the RNG step counter are read with ld.global.cg (L2 only): every element is read once.
__device__ __forceinline__ void head_rng_resolve(RngDev &rng) {
  if (rng.step_dev) rng.step += __ldcg(rng.step_dev);
}
// mu, sigma, eps, w of the 4 weights at linear element e (e % 4 == 0) of sample s
__device__ __forceinline__ void head_quad(const LinArgs &a, int s, int64_t e, bool sample, bool need_sigma, HQuad &q) {
  const float4 m = __ldcg(reinterpret_cast<const float4 *>(a.w_mu + e));
  q.mu[0] = m.x; q.mu[1] = m.y; q.mu[2] = m.z; q.mu[3] = m.w;
  if (sample || need_sigma) {
    const float4 r = __ldcg(reinterpret_cast<const float4 *>(a.w_rho + e));
    q.sg[0] = softplus_f(r.x); q.sg[1] = softplus_f(r.y); q.sg[2] = softplus_f(r.z); q.sg[3] = softplus_f(r.w);
  }
  if (sample) {
    if (a.eps_w) {
      const float4 t = __ldg(reinterpret_cast<const float4 *>(a.eps_w + (int64_t)s * a.out * a.in + e));
      q.ep[0] = t.x; q.ep[1] = t.y; q.ep[2] = t.z; q.ep[3] = t.w;
    } else {
      philox_normal4(a.rng, a.rng.tensor_w, a.rng.sample_base + (uint32_t)s, (uint32_t)(e >> 2), q.ep);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) q.w[j] = __fadd_rn(q.mu[j], __fmul_rn(q.sg[j], q.ep[j]));
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) { q.ep[j] = 0.0f; q.w[j] = q.mu[j]; }
  }
}
__device__ __forceinline__ void head_bias(const LinArgs &a, int s, int64_t o, bool sample, bool need_sigma, float &b,
                                          float &sg, float &ep) {
  const float mu = __ldcg(a.b_mu + o);
  sg = (sample || need_sigma) ? softplus_f(__ldcg(a.b_rho + o)) : 0.0f;
  ep = 0.0f;
  if (sample)
    ep = a.eps_b ? __ldg(a.eps_b + (int64_t)s * a.out + o)
                 : philox_normal1(a.rng, a.rng.tensor_b, a.rng.sample_base + (uint32_t)s, (uint64_t)o);
  b = sample ? __fadd_rn(mu, __fmul_rn(sg, ep)) : mu;
}

// ---- cluster plumbing ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_size() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// the address, in the shared::cluster window, of `p`'s offset inside CTA `rank` of this cluster
__device__ __forceinline__ uint32_t peer_addr(const void *p, uint32_t rank) {
  const uint32_t local = (uint32_t)__cvta_generic_to_shared(p);
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(rank));
  return remote;
}
__device__ __forceinline__ float ld_cluster(uint32_t addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_cluster(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

constexpr int FT = 512;    // threads of the forward
constexpr int SPC = 2;     // most samples per cluster
constexpr int XQ = 10;     // activation quads a thread keeps in flight
constexpr int CLMAX = 16;  // largest cluster

struct HeadFwdArgs {
  int nll_kind;              // BBB_NLL_*
  const int64_t *target_i;   // [B]      (cross-entropy)
  const float *target_f;     // [B, out] (Gaussian)
  float inv_2var, inv_var, cst, grad_scale;
  float *dy;                 // [S, B, out] d nll / d y * grad_scale, nullable
  double *nll;               // += sum over (s, b), nullable when nll_kind == NONE
  float beta;
  const float *beta_dev;
  float *out4;               // nullable: no ELBO assembly
  uint32_t *done;            // zeroed device counter (needed with out4 when there are several clusters)
  int rpc_log2, spc, tpr;    // rows of a sample per cluster (power of two), samples per cluster, lanes per row
  int n_clusters;
};

// ==================================================================================================
// head forward: a cluster owns `spc` samples x `rpc` batch rows; CTA c of the cluster owns the c-th k range
// ==================================================================================================
// NOUT: compiled output width; EXACT: out == NOUT (no column predicates in the unrolled loops)
template <int NOUT, bool EXACT>
__global__ void __launch_bounds__(FT, 1) head_fwd_kernel(const LinArgs a_in, const HeadFwdArgs h, int kqp) {
  extern __shared__ __align__(16) float dyn[];     // Ws [spc][kqp quads][NOUT] float4 | part [R][NO]
  __shared__ float bias_s[SPC][NO];
  __shared__ float red_s[2 * SPC + 1][FT / 32];
  __shared__ float cl_s[CLMAX][8];                 // on rank 0: the (lp, lq per sample, nll) sums of every CTA
  pdl_launch_dependents();
  pdl_wait();
  LinArgs a = a_in;
  head_rng_resolve(a.rng);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cluster_rank(), cln = cluster_size();
  const int cl = blockIdx.x / (int)cln;
  const int rpc = 1 << h.rpc_log2, spc = h.spc, tpr = h.tpr, R = spc * rpc;
  const int bps = (int)((a.B + rpc - 1) >> h.rpc_log2);     // row blocks per sample
  const int sgi = cl / bps, rb = cl - sgi * bps;
  const int s0 = sgi * spc, ns = min(spc, a.S - s0);
  const int64_t row0 = (int64_t)rb << h.rpc_log2;
  const bool sample = a.flags & BBB_F_SAMPLE, relu = a.flags & BBB_F_RELU_IN;
  const bool lpcta = (a.flags & BBB_F_LOGPROB) && rb == 0;   // one cluster per sample group owns the log-prob terms
  const int in = (int)a.in, out = (int)a.out, nq = in >> 2;
  const int q0 = (int)((int64_t)rank * nq / cln), q1 = (int)((int64_t)(rank + 1) * nq / cln), kq = q1 - q0;
  float4 *Ws = reinterpret_cast<float4 *>(dyn);
  float *part = dyn + spc * kqp * NOUT * 4;
  auto col_ok = [&](int o) { return EXACT || o < out; };
  float lpv[SPC] = {0.f, 0.f}, lqv[SPC] = {0.f, 0.f}, nl = 0.0f;

  // thread -> (sample of the cluster, row of the block, lane of the row)
  const int row_all = tid / tpr, sub = tid - row_all * tpr;
  const int sl_t = row_all >> h.rpc_log2, rl_t = row_all & (rpc - 1);
  const bool row_ok = sl_t < ns && row0 + rl_t < a.B;
  const float *xr = a.x + (int64_t)(s0 + sl_t) * a.x_sstride + (row0 + rl_t) * a.in + ((int64_t)q0 << 2);
  float4 xv[XQ];
  auto load_chunk = [&](int c) {
#pragma unroll
    for (int j = 0; j < XQ; ++j) {
      const int q = sub + tpr * (c * XQ + j);
      xv[j] = (row_ok && q < kq) ? __ldg(reinterpret_cast<const float4 *>(xr + 4 * q)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  load_chunk(0);     // in flight while the weights are sampled

  // 1. this CTA's k range of the weights of its samples, each quad formed once in the cluster
  for (int idx = tid; idx < ns * out * kq; idx += FT) {
    const int sl = idx / (out * kq), r = idx - sl * out * kq, o = r / kq, q = r - o * kq;
    HQuad hq;
    head_quad(a, s0 + sl, (int64_t)o * in + ((int64_t)(q0 + q) << 2), sample, lpcta, hq);
    Ws[(sl * kqp + q) * NOUT + o] = make_float4(hq.w[0], hq.w[1], hq.w[2], hq.w[3]);
    if (lpcta) {
      float p = 0.0f, g = 0.0f;
#pragma unroll
      for (int j = 0; j < 4; ++j) { p += logp_elem(a.prior, hq.w[j]); g += logq_elem(hq.sg[j], hq.ep[j]); }
      if (sl == 0) { lpv[0] += p; lqv[0] += g; } else { lpv[1] += p; lqv[1] += g; }
    }
  }
  if (tid < ns * out) {
    const int sl = tid / out, o = tid - sl * out;
    float bv, sg, ep;
    head_bias(a, s0 + sl, o, sample, lpcta, bv, sg, ep);
    bias_s[sl][o] = bv;
    if (lpcta && rank == 0) {
      const float p = logp_elem(a.prior, bv), g = logq_elem(sg, ep);
      if (sl == 0) { lpv[0] += p; lqv[0] += g; } else { lpv[1] += p; lqv[1] += g; }
    }
  }
  __syncthreads();

  // 2. partial dot products over this CTA's k range: `tpr` adjacent lanes share a row and stride over the quads
  {
    float acc[NOUT];
#pragma unroll
    for (int o = 0; o < NOUT; ++o) acc[o] = 0.0f;
    const float4 *wrow = Ws + sl_t * kqp * NOUT;
    for (int c = 0;;) {
#pragma unroll
      for (int j = 0; j < XQ; ++j) {
        const int q = sub + tpr * (c * XQ + j);
        if (row_ok && q < kq) {
          float4 v = xv[j];
          if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
          const float4 *wq = wrow + q * NOUT;
#pragma unroll
          for (int o = 0; o < NOUT; ++o) {
            if (col_ok(o)) {
              const float4 wv = wq[o];
              acc[o] = fmaf(v.x, wv.x, fmaf(v.y, wv.y, fmaf(v.z, wv.z, fmaf(v.w, wv.w, acc[o]))));
            }
          }
        }
      }
      ++c;
      if (c * XQ * tpr >= kq) break;
      load_chunk(c);
    }
#pragma unroll
    for (int o = 0; o < NOUT; ++o) {
      if (col_ok(o)) {
        float v = acc[o];
        for (int d = tpr >> 1; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if (sub == 0) part[row_all * NO + o] = v;
      }
    }
  }
  cluster_arrive();
  cluster_wait();

  // 3. CTA c finishes rows [c R/cln, (c+1) R/cln) of the cluster: 16 lanes per row (one per output column) sum the
  //    k ranges of all CTAs in a fixed order, add the bias, write y and evaluate the likelihood term of the row
  {
    const int rows_own = R / (int)cln, c = tid & 15;
    uint32_t peer[CLMAX];
#pragma unroll
    for (uint32_t p = 0; p < (uint32_t)CLMAX; ++p) peer[p] = p < cln ? peer_addr(part, p) : 0u;
    for (int r0 = 0; r0 < rows_own; r0 += FT / 16) {
      const int ro = r0 + (tid >> 4);
      const int row = (int)rank * rows_own + ro;                 // row of the cluster
      const int sl = row >> h.rpc_log2;
      const int64_t b = row0 + (row & (rpc - 1));
      const bool ok = ro < rows_own && sl < ns && b < a.B;
      const bool col = c < out;
      float v = 0.0f;
      if (ok && col) {
        const uint32_t off = (uint32_t)(row * NO + c) * 4u;
#pragma unroll
        for (uint32_t p = 0; p < (uint32_t)CLMAX; ++p)
          if (p < cln) v += ld_cluster(peer[p] + off);
        v += bias_s[sl][c];
        a.y[((int64_t)(s0 + sl) * a.B + b) * out + c] = v;
      }
      if (h.nll_kind == BBB_NLL_CE) {
        float mx = (ok && col) ? v : -INFINITY;
#pragma unroll
        for (int d = 8; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        float se = (ok && col) ? expf(v - mx) : 0.0f;
#pragma unroll
        for (int d = 8; d > 0; d >>= 1) se += __shfl_xor_sync(0xffffffffu, se, d);
        if (ok && col) {
          const float lse = mx + logf(se);
          const int64_t t = h.target_i[b];
          if (c == t) nl += lse - v;
          if (h.dy) h.dy[((int64_t)(s0 + sl) * a.B + b) * out + c] = h.grad_scale * (expf(v - lse) - (c == t ? 1.0f : 0.0f));
        }
      } else if (h.nll_kind == BBB_NLL_GAUSS) {
        if (ok && col) {
          const float d = v - h.target_f[b * out + c];
          nl += fmaf(d * d, h.inv_2var, h.cst);
          if (h.dy) h.dy[((int64_t)(s0 + sl) * a.B + b) * out + c] = h.grad_scale * d * h.inv_var;
        }
      }
    }
  }

  // 4. CTA sums -> rank 0 of the cluster (distributed shared memory) -> fp64 accumulators and the ELBO scalars
  {
    float vals[2 * SPC + 1] = {lpv[0], lqv[0], lpv[1], lqv[1], nl};
#pragma unroll
    for (int k = 0; k < 2 * SPC + 1; ++k) {
      const float v = warp_sum(vals[k]);
      if (lane == 0) red_s[k][warp] = v;
    }
    __syncthreads();
    if (tid < 2 * SPC + 1) {
      float v = 0.0f;
#pragma unroll
      for (int w = 0; w < FT / 32; ++w) v += red_s[tid][w];
      st_cluster(peer_addr(&cl_s[rank][tid], 0), v);
    }
  }
  cluster_arrive();            // also: this CTA no longer reads its peers' `part`
  cluster_wait();
  if (rank == 0 && tid == 0) {
    double tot[2 * SPC + 1];
#pragma unroll
    for (int k = 0; k < 2 * SPC + 1; ++k) {
      double v = 0.0;
      for (uint32_t p = 0; p < cln; ++p) v += (double)cl_s[p][k];
      tot[k] = v;
    }
    const int S = a.S;
    double lp_new[SPC] = {0.0, 0.0}, lq_new[SPC] = {0.0, 0.0}, nll_new = 0.0;
#pragma unroll
    for (int sl = 0; sl < SPC; ++sl) {
      if (sl < ns) {
        if (lpcta) {
          lp_new[sl] = atomicAdd(a.logp + s0 + sl, tot[2 * sl]) + tot[2 * sl];
          lq_new[sl] = atomicAdd(a.logq + s0 + sl, tot[2 * sl + 1]) + tot[2 * sl + 1];
        } else if (h.out4 && h.n_clusters == 1) {
          lp_new[sl] = __ldcg(a.logp + s0 + sl);
          lq_new[sl] = __ldcg(a.logq + s0 + sl);
        }
      }
    }
    if (h.nll_kind != BBB_NLL_NONE && h.nll) nll_new = atomicAdd(h.nll, tot[2 * SPC]) + tot[2 * SPC];
    if (h.out4) {
      bool last = true;
      if (h.n_clusters > 1) {   // several clusters: the one that arrives last reads the completed accumulators
        __threadfence();
        last = atomicAdd(h.done, 1u) == (uint32_t)h.n_clusters - 1u;
        if (last) {
          __threadfence();
          *h.done = 0u;          // the counter is ready for the next call
        }
      }
      if (last) {
        float beta = h.beta;
        if (h.beta_dev) beta *= __ldg(h.beta_dev);
        double slp = 0.0, slq = 0.0;
        if (h.n_clusters > 1) {
          nll_new = __ldcg(h.nll);
          for (int i = 0; i < S; ++i) { slp += (double)(float)__ldcg(a.logp + i); slq += (double)(float)__ldcg(a.logq + i); }
        } else {                 // the only cluster holds every sample: the values its own atomics produced
#pragma unroll
          for (int i = 0; i < SPC; ++i)
            if (i < S) { slp += (double)(float)lp_new[i]; slq += (double)(float)lq_new[i]; }
        }
        const float nll_m = (float)(nll_new / S);
        const float lpm = (float)(slp / S), lqm = (float)(slq / S);
        h.out4[0] = beta * lqm - beta * lpm + nll_m; h.out4[1] = lpm; h.out4[2] = lqm; h.out4[3] = nll_m;
      }
    }
  }
}

// ==================================================================================================
// head backward
// ==================================================================================================
constexpr int HT = 256;     // threads of the backward
constexpr int RP = 256;     // (sample, batch row) pairs per pass: one per thread
constexpr int CWM = 16;     // widest column range of a CTA
constexpr int SPM = 4;      // most samples per pass
constexpr int DZP = 20;     // shared-memory row pitches (floats): 16-byte aligned rows, conflict-free 128-bit reads
constexpr int XP = 20;
constexpr int SLOTS = 64;   // (sample of the pass, o, column quad) items of a CTA; HT / SLOTS batch groups each
constexpr int NG = HT / SLOTS;
constexpr int kBwdDyn = (RP * DZP + RP * XP) * 4;

template <int NOUT, bool EXACT>
__global__ void __launch_bounds__(HT) head_bwd_kernel(const LinArgs a_in, int cw, int spp, int rbp) {
  extern __shared__ __align__(16) float dynb[];          // dz_s [pair][DZP] | x_s [pair][XP]
  __shared__ __align__(16) float Wt[SPM * NO * CWM];     // [sample of the pass][o][i] sampled weights
  __shared__ __align__(16) float Gp[NG][SLOTS][4];       // per batch group partial sum_b dz x of an item's quad
  __shared__ float Cp[NG][SLOTS];                        // per batch group partial sum_b dz   (bias)
  __shared__ __align__(16) float GM[SLOTS][4], GR[SLOTS][4];   // per item contributions to grad_mu / grad_rho
  __shared__ float BM[SLOTS], BR[SLOTS];
  pdl_launch_dependents();
  pdl_wait();
  LinArgs a = a_in;
  head_rng_resolve(a.rng);
  float *dz_s = dynb, *x_s = dynb + RP * DZP;
  const int tid = threadIdx.x;
  const bool sample = a.flags & BBB_F_SAMPLE, relu = a.flags & BBB_F_RELU_IN, wgrad = !(a.flags & BBB_F_NO_WGRAD);
  const bool want_dx = !(a.flags & BBB_F_NO_DX), dx_preact = a.flags & BBB_F_DX_PREACT, accum = a.flags & BBB_F_ACCUM;
  const int out = EXACT ? NOUT : (int)a.out;
  auto col_ok = [&](int o) { return EXACT || o < out; };
  const int64_t i_lo = (int64_t)blockIdx.x * cw;
  const int w = (int)min((int64_t)cw, a.in - i_lo), nqc = w >> 2;   // columns / quads of this CTA
  const int per = out * nqc;                                         // weight quads of this CTA
  const bool bias_cta = blockIdx.x == 0;

  // item slot -> (sample of the pass, o, column quad); batch group of this thread for the wgrad partial sums
  const int slot = tid & (SLOTS - 1), grp = tid / SLOTS;
  const int it_sl = slot / per, it_r = slot - it_sl * per, it_o = it_r / nqc, it_q = it_r - it_o * nqc;
  const bool ithread = tid < SLOTS;                                  // samples / finishes item `slot`
  const int64_t it_e = (int64_t)it_o * a.in + i_lo + it_q * 4;
  // owner thread t < per accumulates the gradient of weight quad t over all samples
  float gm[4] = {0.f, 0.f, 0.f, 0.f}, gr[4] = {0.f, 0.f, 0.f, 0.f};
  float gbm = 0.0f, gbr = 0.0f;
  HQuad hq;
  auto sample_item = [&](int s0) {
    head_quad(a, s0 + it_sl, it_e, sample, true, hq);
    *reinterpret_cast<float4 *>(&Wt[(it_sl * NO + it_o) * CWM + it_q * 4]) = make_float4(hq.w[0], hq.w[1], hq.w[2], hq.w[3]);
  };
  const float osc = a.out_scale_dev ? __ldg(a.out_scale_dev) : 1.0f;
  const float dxs = ((a.flags & BBB_F_SCALE_DX) && a.out_scale_dev) ? osc : 1.0f;

  for (int s0 = 0; s0 < a.S; s0 += spp) {
    const int ns = min(spp, a.S - s0);
    const int items = ns * per;
    const bool item_ok = slot < items;
    float4 gacc = make_float4(0.f, 0.f, 0.f, 0.f);
    float cacc = 0.0f;
    bool need_sample = true;

    for (int64_t b0 = 0; b0 < a.B; b0 += rbp) {
      const int nb = (int)min((int64_t)rbp, a.B - b0);
      const int pairs = ns * nb;                         // pair p = sl * nb + b, one per thread
      const int p_sl = tid / nb, p_b = tid - p_sl * nb;
      const bool p_ok = tid < pairs;
      // 1. this thread's pair: its x columns and its dz row, straight into registers
      float4 xq[CWM / 4];
      float dzr[NOUT];
      {
        const float *xs = a.x + (int64_t)(s0 + p_sl) * a.x_sstride + (b0 + p_b) * a.in + i_lo;
#pragma unroll
        for (int q = 0; q < CWM / 4; ++q)
          xq[q] = (p_ok && q < nqc) ? __ldg(reinterpret_cast<const float4 *>(xs) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        const int64_t base = ((int64_t)(s0 + p_sl) * a.B + b0 + p_b) * out;
#pragma unroll
        for (int o = 0; o < NOUT; ++o) {
          float v = (p_ok && col_ok(o)) ? __ldg(a.dy + base + o) : 0.0f;
          if (a.mask && p_ok && col_ok(o) && !(__ldg(a.mask + base + o) > 0.0f)) v = 0.0f;
          dzr[o] = v;
        }
      }
      // 2. the item threads sample their weight quad while those loads are in flight
      if (need_sample) {
        need_sample = false;
        if (ithread && item_ok) sample_item(s0);
      }
      if (p_ok) {          // the previous pass was consumed before its barrier (C)
#pragma unroll
        for (int q = 0; q < CWM / 4; ++q) {
          if (relu) { xq[q].x = fmaxf(xq[q].x, 0.f); xq[q].y = fmaxf(xq[q].y, 0.f); xq[q].z = fmaxf(xq[q].z, 0.f); xq[q].w = fmaxf(xq[q].w, 0.f); }
          if (q < nqc) *reinterpret_cast<float4 *>(x_s + tid * XP + 4 * q) = xq[q];
        }
#pragma unroll
        for (int o = 0; o < NOUT; ++o) dz_s[tid * DZP + o] = dzr[o];
      }
      __syncthreads();   // (B) dz_s, x_s, Wt visible
      if (wgrad && item_ok) {      // G_s[o][i quad] partial over this thread's share of the batch rows
        const int bA = (int)((int64_t)grp * nb / NG), bB = (int)((int64_t)(grp + 1) * nb / NG);
        const float *dzp = dz_s + (it_sl * nb) * DZP + it_o, *xp = x_s + (it_sl * nb) * XP + it_q * 4;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        float c = 0.0f;
#pragma unroll 4
        for (int b = bA; b < bB; ++b) {
          const float d = dzp[b * DZP];
          const float4 xv = *reinterpret_cast<const float4 *>(xp + b * XP);
          g.x = fmaf(d, xv.x, g.x); g.y = fmaf(d, xv.y, g.y); g.z = fmaf(d, xv.z, g.z); g.w = fmaf(d, xv.w, g.w);
          c += d;
        }
        *reinterpret_cast<float4 *>(&Gp[grp][slot][0]) = g;
        Cp[grp][slot] = c;
      }
      if (want_dx && p_ok) {       // dX_s[b][i] = sum_o dz_s[b][o] W_s[o][i]: complete, one row per thread
        float *dxr = a.dx + ((int64_t)(s0 + p_sl) * a.B + b0 + p_b) * a.in + i_lo;
        const float *wp = Wt + p_sl * NO * CWM;
#pragma unroll
        for (int q = 0; q < CWM / 4; ++q) {
          if (q < nqc) {
            float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int o = 0; o < NOUT; ++o) {
              if (col_ok(o)) {
                const float4 wv = *reinterpret_cast<const float4 *>(wp + o * CWM + 4 * q);
                d.x = fmaf(dzr[o], wv.x, d.x); d.y = fmaf(dzr[o], wv.y, d.y);
                d.z = fmaf(dzr[o], wv.z, d.z); d.w = fmaf(dzr[o], wv.w, d.w);
              }
            }
            d.x *= dxs; d.y *= dxs; d.z *= dxs; d.w *= dxs;
            if (dx_preact) {
              if (!(xq[q].x > 0.f)) d.x = 0.f;
              if (!(xq[q].y > 0.f)) d.y = 0.f;
              if (!(xq[q].z > 0.f)) d.z = 0.f;
              if (!(xq[q].w > 0.f)) d.w = 0.f;
            }
            *reinterpret_cast<float4 *>(dxr + 4 * q) = d;
          }
        }
      }
      __syncthreads();   // (C) partial sums written; dz_s / x_s / Wt consumed
      if (wgrad && ithread && item_ok) {
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          const float4 v = *reinterpret_cast<const float4 *>(&Gp[g][slot][0]);
          gacc.x += v.x; gacc.y += v.y; gacc.z += v.z; gacc.w += v.w;
          cacc += Cp[g][slot];
        }
      }
    }
    // 3. analytic mu/rho-gradient epilogue of the items, then the owners add up the samples of the pass
    if (wgrad && ithread && item_ok) {
      const int sgl = s0 + it_sl;
      const float gps = a.gp * (a.gp_dev ? __ldg(a.gp_dev + sgl * a.g_dev_stride) : 1.0f);
      const float gqs = a.gq * (a.gq_dev ? __ldg(a.gq_dev + sgl * a.g_dev_stride) : 1.0f);
      const float G[4] = {gacc.x, gacc.y, gacc.z, gacc.w};
      float tm[4], tr[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float t = G[j];
        if (gps != 0.0f) t = fmaf(-gps * hq.w[j], prior_R(a.prior, hq.w[j]), t);
        tm[j] = t;
        tr[j] = -expm1f(-hq.sg[j]) * (t * hq.ep[j] - gqs / hq.sg[j]);   // sigmoid(rho) = 1 - e^-sigma
      }
      *reinterpret_cast<float4 *>(&GM[slot][0]) = make_float4(tm[0], tm[1], tm[2], tm[3]);
      *reinterpret_cast<float4 *>(&GR[slot][0]) = make_float4(tr[0], tr[1], tr[2], tr[3]);
      if (bias_cta && it_q == 0) {   // the item (sl, o, first quad) also carries the bias element o of its sample
        float bv, sg, ep;
        head_bias(a, sgl, it_o, sample, true, bv, sg, ep);
        float t = cacc;
        if (gps != 0.0f) t = fmaf(-gps * bv, prior_R(a.prior, bv), t);
        BM[slot] = t;
        BR[slot] = -expm1f(-sg) * (t * ep - gqs / sg);
      }
    }
    __syncthreads();     // (D)
    if (wgrad && tid < per) {
      for (int sl = 0; sl < ns; ++sl) {
        const float4 m = *reinterpret_cast<const float4 *>(&GM[sl * per + tid][0]);
        const float4 r = *reinterpret_cast<const float4 *>(&GR[sl * per + tid][0]);
        gm[0] += m.x; gm[1] += m.y; gm[2] += m.z; gm[3] += m.w;
        gr[0] += r.x; gr[1] += r.y; gr[2] += r.z; gr[3] += r.w;
      }
      if (bias_cta && tid - (tid / nqc) * nqc == 0) {
        for (int sl = 0; sl < ns; ++sl) { gbm += BM[sl * per + tid]; gbr += BR[sl * per + tid]; }
      }
    }
    // the next pass writes GM / GR / BM / BR only after its barriers (A)-(C)
  }
  if (wgrad && tid < per) {
    const int o = tid / nqc, q = tid - o * nqc;
    const int64_t e = (int64_t)o * a.in + i_lo + q * 4;
    float4 *pm = reinterpret_cast<float4 *>(a.g_w_mu + e), *pr = reinterpret_cast<float4 *>(a.g_w_rho + e);
    float4 om = make_float4(0.f, 0.f, 0.f, 0.f), orr = om;
    if (accum) { om = *pm; orr = *pr; }
    *pm = make_float4(fmaf(osc, gm[0], om.x), fmaf(osc, gm[1], om.y), fmaf(osc, gm[2], om.z), fmaf(osc, gm[3], om.w));
    *pr = make_float4(fmaf(osc, gr[0], orr.x), fmaf(osc, gr[1], orr.y), fmaf(osc, gr[2], orr.z), fmaf(osc, gr[3], orr.w));
    if (bias_cta && q == 0) {
      a.g_b_mu[o] = accum ? fmaf(osc, gbm, a.g_b_mu[o]) : osc * gbm;
      a.g_b_rho[o] = accum ? fmaf(osc, gbr, a.g_b_rho[o]) : osc * gbr;
    }
  }
}

// ==================================================================================================
// local-reparameterisation head backward (weights [in, out], out <= 16; networks.py:116-138 differentiated):
//   dV = dz eps_a / (2 delta);   g1[i][o] = sum_{s,b} x dz;   g2[i][o] = sum_{s,b} x^2 dV
//   grad_mu = g1 + gk mu / sp^2;   grad_rho = sigmoid(rho) (2 sigma g2 + gk (sigma / sp^2 - 1 / sigma))
//   dx[b][i] = sum_o dz mu[i][o] + 2 x sum_o dV sigma^2[i][o]
// Same organisation as head_bwd_kernel: a CTA owns the weight rows i_lo .. i_lo + cw completely, a pass covers up to
// 256 (sample, batch-row) pairs, one per thread, with its dz / dV rows and x columns in registers.  mu and sigma are
// sample-independent, so g1 / g2 simply accumulate over every pass in the registers of their (i, o) item threads; no
// weight is sampled here -- only the activation noise eps_a is regenerated, one Philox call per quad of the
// pair's output row.
// ==================================================================================================
template <int NOUT, bool EXACT>
__global__ void __launch_bounds__(HT) lr_head_bwd_kernel(const LrArgs a_in, int cw, int spp, int rbp) {
  extern __shared__ __align__(16) float dynb[];          // dz_s [pair][DZP] | dv_s [pair][DZP] | x_s [pair][XP]
  __shared__ float Mu[CWM][NO], S2[CWM][NO];             // this CTA's rows of mu and sigma^2
  __shared__ float Gp[2][NG][CWM * NO];                  // per batch group partial g1 / g2 of a pass
  __shared__ float Cp[SPM][NO];                          // per sample of the pass: sum_b dz  (bias gradients)
  pdl_launch_dependents();
  pdl_wait();
  LrArgs a = a_in;
  head_rng_resolve(a.rng);
  float *dz_s = dynb, *dv_s = dynb + RP * DZP, *x_s = dynb + 2 * RP * DZP;
  const int tid = threadIdx.x;
  const bool sample = a.flags & BBB_F_SAMPLE, relu = a.flags & BBB_F_RELU_IN, wgrad = !(a.flags & BBB_F_NO_WGRAD);
  const bool want_dx = !(a.flags & BBB_F_NO_DX), dx_preact = a.flags & BBB_F_DX_PREACT, accum = a.flags & BBB_F_ACCUM;
  const int out = EXACT ? NOUT : (int)a.out;
  auto col_ok = [&](int o) { return EXACT || o < out; };
  const int64_t i_lo = (int64_t)blockIdx.x * cw;
  const int w = (int)min((int64_t)cw, a.in - i_lo), nqc = w >> 2;
  const bool bias_cta = blockIdx.x == 0;
  const float osc = a.out_scale_dev ? __ldg(a.out_scale_dev) : 1.0f;
  const float dxs = ((a.flags & BBB_F_SCALE_DX) && a.out_scale_dev) ? osc : 1.0f;
  const float gk = (a.flags & BBB_F_LOGPROB) ? a.g_kl * (a.g_kl_dev ? __ldg(a.g_kl_dev) : 1.0f) : 0.0f;
  const float inv_sp2 = 1.0f / (a.sigma_p * a.sigma_p);

  // item thread t < w * out owns weight (i_lo + t / out, t % out): mu, sigma once; g1, g2 over all passes
  const bool ithread = tid < w * out;
  const int wi = ithread ? tid / out : 0, wo = ithread ? tid - wi * out : 0;
  const int64_t we = (i_lo + wi) * a.out + wo;
  float mu = 0.0f, sg = 1.0f, g1 = 0.0f, g2 = 0.0f, gbmu = 0.0f, gbrho = 0.0f;
  if (ithread) {
    mu = __ldcg(a.w_mu + we);
    sg = softplus_f(__ldcg(a.w_rho + we));
    Mu[wi][wo] = mu;
    S2[wi][wo] = sg * sg;
  }
  const int items = w * out;
  const int it = tid % 64, grp = tid / 64;                // 64 item lanes x 4 batch groups; items beyond 64: second round

  for (int s0 = 0; s0 < a.S; s0 += spp) {
    const int ns = min(spp, a.S - s0);
    if (tid < SPM * NO) Cp[tid / NO][tid % NO] = 0.0f;
    float p1[(CWM * NO + 63) / 64], p2[(CWM * NO + 63) / 64];
#pragma unroll
    for (int r = 0; r < (CWM * NO + 63) / 64; ++r) p1[r] = p2[r] = 0.0f;
    for (int64_t b0 = 0; b0 < a.B; b0 += rbp) {
      const int nb = (int)min((int64_t)rbp, a.B - b0);
      const int pairs = ns * nb;
      const int p_sl = tid / nb, p_b = tid - p_sl * nb;
      const bool p_ok = tid < pairs;
      float4 xq[CWM / 4];
      float dzr[NOUT], dvr[NOUT];
      {
        const float *xs = a.x + (int64_t)(s0 + p_sl) * a.x_sstride + (b0 + p_b) * a.in + i_lo;
#pragma unroll
        for (int q = 0; q < CWM / 4; ++q)
          xq[q] = (p_ok && q < nqc) ? __ldg(reinterpret_cast<const float4 *>(xs) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        const int s = s0 + p_sl;
        const int64_t row = (int64_t)(b0 + p_b) * out;                 // index of (b, o = 0) inside the sample
        const int64_t base = (int64_t)s * a.B * a.out + row;
        float dl[NOUT];
#pragma unroll
        for (int o = 0; o < NOUT; ++o) {
          float z = (p_ok && col_ok(o)) ? __ldg(a.dy + base + o) : 0.0f;
          if (a.mask && p_ok && col_ok(o) && !(__ldg(a.mask + base + o) > 0.0f)) z = 0.0f;
          dzr[o] = z;
          dl[o] = (sample && p_ok && col_ok(o)) ? __ldg(a.delta_in + base + o) : 0.0f;
          dvr[o] = 0.0f;
        }
        if (sample && p_ok) {
          if (a.eps_a) {
#pragma unroll
            for (int o = 0; o < NOUT; ++o)
              if (col_ok(o) && dl[o] > 0.0f) dvr[o] = dzr[o] * __ldg(a.eps_a + base + o) / (2.0f * dl[o]);
          } else {
            // eps_a of elements row .. row + out - 1: the Philox quads that cover them, one call each
            const uint32_t smp = a.rng.sample_base + (uint32_t)s;
            int64_t q4 = row >> 2;
            float e4[4];
            philox_normal4(a.rng, a.rng.tensor_w, smp, (uint32_t)q4, e4);
#pragma unroll
            for (int o = 0; o < NOUT; ++o) {
              if (col_ok(o)) {
                const int64_t idx = row + o;
                if ((idx >> 2) != q4) { q4 = idx >> 2; philox_normal4(a.rng, a.rng.tensor_w, smp, (uint32_t)q4, e4); }
                const int l = (int)(idx & 3);
                const float ep = l == 0 ? e4[0] : l == 1 ? e4[1] : l == 2 ? e4[2] : e4[3];
                if (dl[o] > 0.0f) dvr[o] = dzr[o] * ep / (2.0f * dl[o]);
              }
            }
          }
        }
      }
      if (p_ok) {
#pragma unroll
        for (int q = 0; q < CWM / 4; ++q) {
          if (relu) { xq[q].x = fmaxf(xq[q].x, 0.f); xq[q].y = fmaxf(xq[q].y, 0.f); xq[q].z = fmaxf(xq[q].z, 0.f); xq[q].w = fmaxf(xq[q].w, 0.f); }
          if (q < nqc) *reinterpret_cast<float4 *>(x_s + tid * XP + 4 * q) = xq[q];
        }
#pragma unroll
        for (int o = 0; o < NOUT; ++o) { dz_s[tid * DZP + o] = dzr[o]; dv_s[tid * DZP + o] = dvr[o]; }
      }
      __syncthreads();   // (B) dz_s, dv_s, x_s (and Mu / S2 on the first pass) visible
      if (wgrad) {
        // g1 / g2 partial sums: item = (i, o) in rounds of 64, this thread's quarter of the pairs
        const int pA = (int)((int64_t)grp * pairs / NG), pB = (int)((int64_t)(grp + 1) * pairs / NG);
#pragma unroll
        for (int r = 0; r < (CWM * NO + 63) / 64; ++r) {
          const int item = it + 64 * r;
          if (item < items) {
            const int i = item / out, o = item - i * out;
            float s1 = 0.0f, s2 = 0.0f;
#pragma unroll 4
            for (int p = pA; p < pB; ++p) {
              const float x = x_s[p * XP + i];
              s1 = fmaf(x, dz_s[p * DZP + o], s1);
              s2 = fmaf(x * x, dv_s[p * DZP + o], s2);
            }
            p1[r] += s1; p2[r] += s2;
          }
        }
        if (bias_cta && tid < ns * out) {
          const int sl = tid / out, o = tid - sl * out;
          float c = 0.0f;
          for (int b = 0; b < nb; ++b) c += dz_s[(sl * nb + b) * DZP + o];
          Cp[sl][o] += c;
        }
      }
      if (want_dx && p_ok) {       // one row of dx per thread, its columns complete
        float *dxr = a.dx + ((int64_t)(s0 + p_sl) * a.B + b0 + p_b) * a.in + i_lo;
#pragma unroll
        for (int q = 0; q < CWM / 4; ++q) {
          if (q < nqc) {
            const float xv[4] = {xq[q].x, xq[q].y, xq[q].z, xq[q].w};
            float d[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              float t1 = 0.0f, t2 = 0.0f;
#pragma unroll
              for (int o = 0; o < NOUT; ++o) {
                if (col_ok(o)) {
                  t1 = fmaf(dzr[o], Mu[4 * q + c][o], t1);
                  t2 = fmaf(dvr[o], S2[4 * q + c][o], t2);
                }
              }
              d[c] = dxs * fmaf(2.0f * xv[c], t2, t1);
              if (dx_preact && !(xv[c] > 0.0f)) d[c] = 0.0f;
            }
            *reinterpret_cast<float4 *>(dxr + 4 * q) = make_float4(d[0], d[1], d[2], d[3]);
          }
        }
      }
      __syncthreads();   // (C) this pass's shared tiles are consumed
    }
    // bias gradients of the pass's samples (CTA 0): sum_b dz, weighted by eps_b for rho
    if (wgrad && bias_cta && tid < out) {
      for (int sl = 0; sl < ns; ++sl) {
        const float c = Cp[sl][tid];
        gbmu += c;
        if (sample) {
          const int s = s0 + sl;
          const float eb = a.eps_b ? __ldg(a.eps_b + (int64_t)s * a.out + tid)
                                   : philox_normal1(a.rng, a.rng.tensor_b, a.rng.sample_base + (uint32_t)s, (uint64_t)tid);
          gbrho += c * eb;
        }
      }
    }
    // fold the pass's partial sums of the 4 batch groups into the item threads
    if (wgrad) {
#pragma unroll
      for (int r = 0; r < (CWM * NO + 63) / 64; ++r) {
        const int item = it + 64 * r;
        if (item < items) { Gp[0][grp][item] = p1[r]; Gp[1][grp][item] = p2[r]; }
      }
    }
    __syncthreads();
    if (wgrad && ithread) {
#pragma unroll
      for (int g = 0; g < NG; ++g) { g1 += Gp[0][g][tid]; g2 += Gp[1][g][tid]; }
    }
    __syncthreads();     // Gp / Cp are rewritten by the next pass
  }
  if (wgrad && ithread) {
    const float gm = fmaf(gk * mu, inv_sp2, g1);
    const float gr = -expm1f(-sg) * (2.0f * sg * g2 + gk * (sg * inv_sp2 - 1.0f / sg));
    a.g_w_mu[we] = accum ? fmaf(osc, gm, a.g_w_mu[we]) : osc * gm;
    a.g_w_rho[we] = accum ? fmaf(osc, gr, a.g_w_rho[we]) : osc * gr;
  }
  if (wgrad && bias_cta && tid < out) {
    const float bmu = __ldcg(a.b_mu + tid), bsg = softplus_f(__ldcg(a.b_rho + tid));
    const float gm = fmaf(gk * bmu, inv_sp2, gbmu);
    const float gr = -expm1f(-bsg) * (gbrho + gk * (bsg * inv_sp2 - 1.0f / bsg));
    a.g_b_mu[tid] = accum ? fmaf(osc, gm, a.g_b_mu[tid]) : osc * gm;
    a.g_b_rho[tid] = accum ? fmaf(osc, gr, a.g_b_rho[tid]) : osc * gr;
  }
}

constexpr int kLrBwdDyn = (2 * RP * DZP + RP * XP) * 4;

template <int NOUT, bool EXACT>
int launch_lr_head_bwd_t(const LrArgs &a, int grid, int cw, int spp, int rbp, cudaStream_t st) {
  BBB_CHECK_CUDA(cudaFuncSetAttribute(lr_head_bwd_kernel<NOUT, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLrBwdDyn));
  BBB_CHECK_CUDA(launch_pdl(lr_head_bwd_kernel<NOUT, EXACT>, dim3(grid), dim3(HT), kLrBwdDyn, st, a, cw, spp, rbp));
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}

inline int cdiv_i(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
inline bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

constexpr int kFwdDynMax = 112 * 1024;

// output widths the kernels are compiled for: exact 1, 4, 10 (MNIST), 16; 2-3 run the 4-wide kernel and 5-9, 11-15
// the 16-wide one with column predicates
#define BBB_HEAD_DISPATCH(out, CALL)                   \
  do {                                                 \
    if ((out) == 1) { CALL(1, true); }                 \
    else if ((out) == 4) { CALL(4, true); }            \
    else if ((out) < 4) { CALL(4, false); }            \
    else if ((out) == 10) { CALL(10, true); }          \
    else if ((out) == 16) { CALL(16, true); }          \
    else { CALL(16, false); }                          \
  } while (0)

// 16-CTA clusters (non-portable size) when the device can co-schedule one, else the portable 8
template <int NOUT, bool EXACT>
int head_cluster_size() {
  static int cached = 0;
  if (cached) return cached;
  auto kernel = head_fwd_kernel<NOUT, EXACT>;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdDynMax);
  int cl = 8;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(16);
    cfg.blockDim = dim3(FT);
    cfg.dynamicSmemBytes = 48 * 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 16; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) == cudaSuccess && n >= 1) cl = 16;
  }
  cudaGetLastError();   // a failed query is not an error of the caller's launch
  cached = cl;
  return cl;
}

template <int NOUT, bool EXACT>
int launch_head_fwd_t(const LinArgs &a, HeadFwdArgs h, cudaStream_t st) {
  int cl = head_cluster_size<NOUT, EXACT>();
  // rows of one sample per cluster: the batch rounded up to a power of two in [16, 512]; as many samples (<= 2) as
  // still give every (sample, row) at least one of the 512 threads; the remaining factor is lanes per row
  int rl2 = 4;
  while ((1 << rl2) < a.B && rl2 < 9) ++rl2;
  h.rpc_log2 = rl2;
  const int rpc = 1 << rl2;
  h.spc = (a.S >= 2 && 2 * rpc <= FT) ? 2 : 1;
  h.tpr = FT / (h.spc * rpc);
  const int nq = (int)(a.in >> 2);
  if (nq < cl) cl = 8;                       // fewer quads than CTAs: the smaller cluster (empty k ranges are fine)
  const int kqp = cdiv_i(nq, cl);            // quads of the widest k range
  const size_t smem = ((size_t)h.spc * kqp * NOUT * 4 + (size_t)h.spc * rpc * NO) * sizeof(float);
  if (smem > (size_t)kFwdDynMax) return fail(BBB_EUNSUPPORTED, "bbb_head_fwd: layer too wide for the shared-memory tile");
  h.n_clusters = cdiv_i(a.S, h.spc) * cdiv_i(a.B, rpc);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(cl * h.n_clusters));
  cfg.blockDim = dim3(FT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  BBB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, head_fwd_kernel<NOUT, EXACT>, a, h, kqp));
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}

int launch_head_fwd(const LinArgs &a, const HeadFwdArgs &h, cudaStream_t st) {
#define CALL(N, E) return launch_head_fwd_t<N, E>(a, h, st)
  BBB_HEAD_DISPATCH(a.out, CALL);
#undef CALL
}

template <int NOUT, bool EXACT>
int launch_head_bwd_t(const LinArgs &a, int grid, int cw, int spp, int rbp, cudaStream_t st) {
  BBB_CHECK_CUDA(cudaFuncSetAttribute(head_bwd_kernel<NOUT, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdDyn));
  BBB_CHECK_CUDA(launch_pdl(head_bwd_kernel<NOUT, EXACT>, dim3(grid), dim3(HT), kBwdDyn, st, a, cw, spp, rbp));
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}

}  // namespace

bool head_supported(const LinArgs &a) {
  // rows of 16-byte multiples (vector loads, aligned Philox quads), at most 2^31 rows * columns per sample
  return a.out >= 1 && a.out <= NO && a.vec_in && a.in >= 4 && a.B >= 1 && a.S >= 1 &&
         a.in <= 8 * 1024 && a.B * a.S <= (int64_t)1 << 24;
}

int launch_linear_fwd_head(const LinArgs &a, cudaStream_t st) {
  HeadFwdArgs h{};
  h.nll_kind = BBB_NLL_NONE;
  return launch_head_fwd(a, h, st);
}

int launch_linear_bwd_head(const LinArgs &a, cudaStream_t st) {
  const int nq_i = (int)(a.in / 4);
  // samples per pass: as many as fit 256 (sample, row) pairs (one load phase for all of them), else one sample in
  // chunks of 256 rows
  int spp = 1, rbp = RP;
  if (a.B <= RP / 2) {
    spp = (int)(RP / a.B);
    if (spp > SPM) spp = SPM;
    if (spp > a.S) spp = a.S;
    rbp = (int)a.B;
  }
  // about one column range per SM; the (sample, o, quad) items of a pass must fit the 64 item slots
  int wq = cdiv_i(nq_i, sm_count());
  int cap = SLOTS / (spp * (int)a.out);
  while (cap < 1) { --spp; cap = SLOTS / (spp * (int)a.out); }   // out = 16 with 4 samples: fewer samples per pass
  if (wq > cap) wq = cap;
  if (wq > CWM / 4) wq = CWM / 4;
  if (wq < 1) wq = 1;
#define CALL(N, E) return launch_head_bwd_t<N, E>(a, cdiv_i(nq_i, wq), wq * 4, spp, rbp, st)
  BBB_HEAD_DISPATCH(a.out, CALL);
#undef CALL
}

bool lr_head_bwd_supported(const LrArgs &a) {
  return a.out >= 1 && a.out <= NO && a.vec_in && a.in >= 4 && a.B >= 1 && a.S >= 1 && a.S <= 65535;
}

int launch_lr_bwd_head(const LrArgs &a, cudaStream_t st) {
  const int nq_i = (int)(a.in / 4);
  int spp = 1, rbp = RP;
  if (a.B <= RP / 2) {
    spp = (int)(RP / a.B);
    if (spp > SPM) spp = SPM;
    if (spp > a.S) spp = a.S;
    rbp = (int)a.B;
  }
  int wq = cdiv_i(nq_i, sm_count());               // about one range of weight rows per SM; w * out <= 256 item threads
  int cap = HT / (4 * (int)a.out);
  if (cap < 1) cap = 1;
  if (wq > cap) wq = cap;
  if (wq > CWM / 4) wq = CWM / 4;
  if (wq < 1) wq = 1;
#define CALL(N, E) return launch_lr_head_bwd_t<N, E>(a, cdiv_i(nq_i, wq), wq * 4, spp, rbp, st)
  BBB_HEAD_DISPATCH(a.out, CALL);
#undef CALL
}

}  // namespace bbb

using namespace bbb;

extern "C" int bbb_head_fwd(const float *x, int64_t x_sample_stride, const float *w_mu, const float *w_rho,
                            const float *b_mu, const float *b_rho, const float *eps_w, const float *eps_b,
                            const bbb_rng *rng, const bbb_prior *prior, int64_t S, int64_t B, int64_t in, int64_t out,
                            int32_t flags, int32_t nll_kind, const void *target, float sigma, float grad_scale,
                            float *y, float *dy, double *logp, double *logq, double *nll, float beta,
                            const float *beta_dev, float *out4, uint32_t *done_counter, void *stream) {
  BBB_CHECK_ARG(S >= 1 && B >= 1 && in >= 1 && out >= 1 && S <= 65535, "bad shape");
  BBB_CHECK_ARG(x && w_mu && b_mu && y, "null pointer");
  BBB_CHECK_ARG(x_sample_stride == 0 || x_sample_stride == B * in, "x_sample_stride must be 0 or B*in");
  const bool sample = flags & BBB_F_SAMPLE, lpq = flags & BBB_F_LOGPROB;
  BBB_CHECK_ARG(!(sample || lpq) || (w_rho && b_rho), "rho pointers required");
  BBB_CHECK_ARG(!lpq || (prior && logp && logq), "log-prob outputs and prior required with BBB_F_LOGPROB");
  BBB_CHECK_ARG(!sample || (eps_w && eps_b) || (!eps_w && !eps_b && rng), "give both eps pointers or an rng");
  BBB_CHECK_ARG(!prior || prior->kind == BBB_PRIOR_GAUSSIAN || prior->kind == BBB_PRIOR_MIXTURE, "bad prior kind");
  BBB_CHECK_ARG(nll_kind == BBB_NLL_NONE || nll_kind == BBB_NLL_CE || nll_kind == BBB_NLL_GAUSS, "bad nll_kind");
  BBB_CHECK_ARG(nll_kind == BBB_NLL_NONE || (target && nll), "target and nll accumulator required");
  BBB_CHECK_ARG(nll_kind != BBB_NLL_GAUSS || sigma > 0, "sigma must be positive");
  BBB_CHECK_ARG(!out4 || (done_counter && logp && logq && nll && nll_kind != BBB_NLL_NONE),
                "ELBO assembly needs done_counter, logp, logq and a likelihood");
  LinArgs a{};
  a.x = x; a.x_sstride = x_sample_stride;
  a.w_mu = w_mu; a.w_rho = w_rho; a.b_mu = b_mu; a.b_rho = b_rho; a.eps_w = eps_w; a.eps_b = eps_b;
  a.rng = make_rng_dev(rng);
  if (prior) a.prior = make_prior_dev(prior);
  a.S = (int)S; a.B = B; a.in = in; a.out = out; a.flags = flags;
  a.y = y; a.logp = logp; a.logq = logq;
  a.vec_in = (in % 4 == 0) && al16(x) && al16(w_mu) && al16(w_rho) && al16(eps_w);
  a.vec_out = false;
  if (!head_supported(a))
    return fail(BBB_EUNSUPPORTED, "bbb_head_fwd: needs out <= 16, in a multiple of 4 (<= 8192) and 16-byte aligned rows");
  HeadFwdArgs h{};
  h.nll_kind = nll_kind;
  if (nll_kind == BBB_NLL_CE) h.target_i = static_cast<const int64_t *>(target);
  if (nll_kind == BBB_NLL_GAUSS) {
    h.target_f = static_cast<const float *>(target);
    const double var = (double)sigma * sigma;
    h.inv_2var = (float)(0.5 / var); h.inv_var = (float)(1.0 / var);
    h.cst = (float)(log((double)sigma) + 0.918938533204672741780329736406);
  }
  h.grad_scale = grad_scale; h.dy = dy; h.nll = nll;
  h.beta = beta; h.beta_dev = beta_dev; h.out4 = out4; h.done = done_counter;
  return launch_head_fwd(a, h, (cudaStream_t)stream);
}
