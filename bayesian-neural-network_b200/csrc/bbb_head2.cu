// The network's head (last layer, out <= 16) for the network-level step -- round-2 organisation.
//
// The head has a few thousand weights; what it costs is the length of its chain of dependent phases, and on how many
// SMs they run.  bbb_head.cu runs the forward as ONE 16-CTA cluster (16 SMs busy for ~12 us) and the backward as 100
// column-owning CTAs.  Here both directions run on a full single-wave grid with the work cut the natural way for
// each phase:
//
//   forward  (head2_fwd_kernel)
//     phase 1  every weight quad is sampled ONCE, grid-wide (thread <-> quad): W_s = mu + sigma eps and the log-prior /
//              log-posterior terms; W_s goes to a [S][out][in] scratch in global memory (it stays in L2), the sampled
//              bias to [S][16].
//     grid barrier (one counter in the zero-filled workspace; the grid is one resident wave)
//     phase 2  row-parallel: a CTA takes (sample, batch-row) pairs; 256 threads split K, each forms its partial dot
//              products against the out weight rows, block reduction, then bias, the likelihood term and its gradient
//              d nll / d y for the row.  The CTA that finishes last assembles the four ELBO scalars.
//   backward (head2_bwd_kernel), no eps regeneration for dgrad: it reuses the scratch W_s
//     part A   row-parallel dgrad: dz_prev[s][b][k] = (sum_o dy[s][b][o] W_s[o][k]) (x[s][b][k] > 0), plain stores (the
//              head is the only producer of the last hidden layer's gradient)
//     part B   column-parallel wgrad: a CTA owns a few input columns; dy and its x columns are staged in shared
//              memory once, thread (o, column) walks the batch, then the analytic epilogue per weight with eps
//              regenerated from the Philox counter; CTA 0 adds the bias gradients.
// Exact fp32 (the head is never a tensor-core contraction), deterministic except for the fp64 atomics of the scalars.
#include "bbb_common.cuh"
#include "bbb_kernels.h"
#include "bbb_mlp.h"

namespace bbb {
namespace {

constexpr int HT = 256;            // threads per CTA
constexpr int MAXO = 16;           // widest head

struct Head2Args {
  const float *x;                  // [S][B][in]  input: pre-activation of the layer below (relu_in) or an activation
  const float *w_mu, *w_rho, *b_mu, *b_rho, *eps_w, *eps_b;
  float *ws;                       // [S][out][in] sampled weights, then [S][16] sampled biases (scratch)
  RngDev rng;
  PriorDev prior;
  int S, B, in, out, flags, relu_in;
  // forward
  int nll_kind;
  const int64_t *target_i;
  const float *target_f;
  float inv_2var, inv_var, cst, grad_scale;
  float *y, *dy;
  double *logp, *logq, *nll;
  float beta;
  const float *beta_dev;
  float *out4;
  uint32_t *done;                  // [0] finished-CTA counter, [1] grid barrier
  // backward
  const float *dy_in;
  float gp, gq;
  const float *gp_dev, *gq_dev, *out_scale_dev;
  int g_dev_stride;
  float *dz_prev, *g_w_mu, *g_w_rho, *g_b_mu, *g_b_rho;
};

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float4 ldcg4(const float *p) { return __ldcg(reinterpret_cast<const float4 *>(p)); }

// eps of weight quad q (elements 4q..4q+3 of the [out][in] matrix) for sample s
__device__ __forceinline__ void eps_quad(const Head2Args &a, const RngDev &rng, int s, int64_t q, bool sample, float ep[4]) {
  if (!sample) { ep[0] = ep[1] = ep[2] = ep[3] = 0.0f; return; }
  if (a.eps_w) {
    const float4 t = ldg4(a.eps_w + ((int64_t)s * a.out * a.in + 4 * q));
    ep[0] = t.x; ep[1] = t.y; ep[2] = t.z; ep[3] = t.w;
  } else {
    philox_normal4(rng, rng.tensor_w, rng.sample_base + (uint32_t)s, (uint32_t)q, ep);
  }
}

template <int OUT>
__global__ void __launch_bounds__(HT) head2_fwd_kernel(const Head2Args a) {
  __shared__ float red[2 * 32];
  __shared__ float part[8][MAXO];
  __shared__ float yrow[MAXO];
  __shared__ uint32_t last_s;
  pdl_wait();
  RngDev rng = a.rng;
  rng_resolve(rng);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool sample = a.flags & BBB_F_SAMPLE, lpq = a.flags & BBB_F_LOGPROB;
  const int64_t nq = (int64_t)a.out * a.in / 4;
  float *bs = a.ws + (int64_t)a.S * a.out * a.in;

  // ---- phase 1: sample every weight quad once, grid-wide --------------------------------------------------------------
  for (int s = 0; s < a.S; ++s) {
    float lp = 0.0f, lq = 0.0f;
    for (int64_t q = (int64_t)blockIdx.x * HT + tid; q < nq; q += (int64_t)gridDim.x * HT) {
      const float4 m4 = ldg4(a.w_mu + 4 * q);
      const float mu[4] = {m4.x, m4.y, m4.z, m4.w};
      float sg[4] = {0.f, 0.f, 0.f, 0.f}, ep[4], w[4];
      if (sample || lpq) {
        const float4 r4 = ldg4(a.w_rho + 4 * q);
        sg[0] = softplus_f(r4.x); sg[1] = softplus_f(r4.y); sg[2] = softplus_f(r4.z); sg[3] = softplus_f(r4.w);
      }
      eps_quad(a, rng, s, q, sample, ep);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        w[e] = sample ? __fadd_rn(mu[e], __fmul_rn(sg[e], ep[e])) : mu[e];
        if (lpq) { lp += logp_elem(a.prior, w[e]); lq += logq_elem(sg[e], ep[e]); }
      }
      *reinterpret_cast<float4 *>(a.ws + (int64_t)s * a.out * a.in + 4 * q) = make_float4(w[0], w[1], w[2], w[3]);
    }
    if (blockIdx.x == 0 && tid < a.out) {          // the sampled bias (its log-prob terms are counted here, once)
      const float bmu = __ldg(a.b_mu + tid);
      const float bsg = (sample || lpq) ? softplus_f(__ldg(a.b_rho + tid)) : 0.0f;
      float ep = 0.0f;
      if (sample)
        ep = a.eps_b ? __ldg(a.eps_b + (int64_t)s * a.out + tid)
                     : philox_normal1(rng, rng.tensor_b, rng.sample_base + (uint32_t)s, (uint64_t)tid);
      const float bv = sample ? __fadd_rn(bmu, __fmul_rn(bsg, ep)) : bmu;
      bs[s * MAXO + tid] = bv;
      if (lpq) { lp += logp_elem(a.prior, bv); lq += logq_elem(bsg, ep); }
    }
    if (lpq) block_sum2_atomic(lp, lq, red, a.logp + s, a.logq + s);
  }
  // ---- grid barrier: all sampled weights are in memory (one resident wave; bounded spin) ----------------------------------
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    atomicAdd(a.done + 1, 1u);
    uint32_t seen = 0;
    for (int spin = 0; spin < (1 << 24); ++spin) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(a.done + 1) : "memory");
      if (seen >= gridDim.x) break;
      __nanosleep(32);
    }
  }
  __syncthreads();
  pdl_launch_dependents();

  // ---- phase 2: one (sample, batch row) pair at a time; the threads split K --------------------------------------------
  double nll_cta = 0.0;
  const int rows = a.S * a.B, nkq = a.in >> 2;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const int s = r / a.B, b = r - s * a.B;
    const float *xr = a.x + (int64_t)r * a.in;
    const float *wsr = a.ws + (int64_t)s * a.out * a.in;
    float acc[OUT];
#pragma unroll
    for (int o = 0; o < OUT; ++o) acc[o] = 0.0f;
    for (int kq = tid; kq < nkq; kq += HT) {
      float4 xv = ldcg4(xr + 4 * kq);
      if (a.relu_in) { xv.x = fmaxf(xv.x, 0.f); xv.y = fmaxf(xv.y, 0.f); xv.z = fmaxf(xv.z, 0.f); xv.w = fmaxf(xv.w, 0.f); }
#pragma unroll
      for (int o = 0; o < OUT; ++o) {
        const float4 wv = ldcg4(wsr + (int64_t)o * a.in + 4 * kq);
        acc[o] = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, acc[o]))));
      }
    }
#pragma unroll
    for (int o = 0; o < OUT; ++o) {
      const float v = warp_sum(acc[o]);
      if (lane == 0) part[warp][o] = v;
    }
    __syncthreads();
    if (tid < OUT) {
      float v = __ldcg(bs + s * MAXO + tid);
#pragma unroll
      for (int w8 = 0; w8 < HT / 32; ++w8) v += part[w8][tid];
      yrow[tid] = v;
      a.y[(int64_t)r * OUT + tid] = v;
    }
    __syncthreads();
    if (tid == 0 && a.nll_kind != BBB_NLL_NONE) {
      float *dyr = a.dy ? a.dy + (int64_t)r * OUT : nullptr;
      if (a.nll_kind == BBB_NLL_CE) {       // CrossEntropyLoss(reduction='sum') (networks.py:187)
        const int t = (int)a.target_i[b];
        float mx = yrow[0];
#pragma unroll
        for (int o = 1; o < OUT; ++o) mx = fmaxf(mx, yrow[o]);
        float e[OUT], sum = 0.0f;
#pragma unroll
        for (int o = 0; o < OUT; ++o) { e[o] = expf(yrow[o] - mx); sum += e[o]; }
        const float lse = mx + logf(sum);
        float yt = 0.0f;
#pragma unroll
        for (int o = 0; o < OUT; ++o) {
          if (o == t) yt = yrow[o];
          if (dyr) dyr[o] = a.grad_scale * (e[o] / sum - (o == t ? 1.0f : 0.0f));
        }
        nll_cta += (double)(lse - yt);
      } else {                               // -Normal(y, sigma).log_prob(target).sum() (networks.py:185)
#pragma unroll
        for (int o = 0; o < OUT; ++o) {
          const float d = yrow[o] - a.target_f[(int64_t)b * OUT + o];
          nll_cta += (double)(d * d * a.inv_2var + a.cst);
          if (dyr) dyr[o] = a.grad_scale * d * a.inv_var;
        }
      }
    }
    __syncthreads();
  }
  // ---- the CTA that finishes last assembles the ELBO scalars (as bbb_elbo_finalize does) ------------------------------------
  if (tid == 0) {
    if (a.nll_kind != BBB_NLL_NONE && nll_cta != 0.0) atomicAdd(a.nll, nll_cta);
    __threadfence();
    last_s = atomicAdd(a.done, 1u) == gridDim.x - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (last_s && tid == 0) {
    __threadfence();
    if (a.out4) {
      const float beta = a.beta * (a.beta_dev ? __ldg(a.beta_dev) : 1.0f);
      const volatile double *lp = a.logp, *lq = a.logq, *nl = a.nll;
      double slp = 0.0, slq = 0.0;
      for (int s = 0; s < a.S; ++s) { slp += (double)(float)lp[s]; slq += (double)(float)lq[s]; }
      const float nll_m = (float)(nl[0] / a.S), lpm = (float)(slp / a.S), lqm = (float)(slq / a.S);
      a.out4[0] = beta * lqm - beta * lpm + nll_m; a.out4[1] = lpm; a.out4[2] = lqm; a.out4[3] = nll_m;
    }
    a.done[0] = 0u;
    a.done[1] = 0u;
  }
}

// columns a CTA owns in part B (wgrad): at most KC, a multiple of 4
constexpr int KC = 12;

template <int OUT>
__global__ void __launch_bounds__(HT) head2_bwd_kernel(const Head2Args a) {
  __shared__ float dy_s[HT][MAXO + 1];          // (s, b) pairs of one pass x outputs
  __shared__ float x_s[HT][KC + 1];             // (s, b) pairs x the CTA's columns (post-ReLU)
  __shared__ float g_s[2][MAXO][KC];            // wgrad of up to two samples per pass group
  __shared__ float dyrow[MAXO];
  pdl_wait();
  RngDev rng = a.rng;
  rng_resolve(rng);
  const int tid = threadIdx.x;
  const bool sample = a.flags & BBB_F_SAMPLE;
  const int rows = a.S * a.B, nkq = a.in >> 2;
  const float osc = a.out_scale_dev ? __ldg(a.out_scale_dev) : 1.0f;

  // ---- part A: dgrad, one (sample, batch row) pair at a time ------------------------------------------------------------
  if (a.dz_prev) {
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
      const int s = r / a.B;
      __syncthreads();
      if (tid < OUT) dyrow[tid] = __ldg(a.dy_in + (int64_t)r * OUT + tid);
      __syncthreads();
      const float *wsr = a.ws + (int64_t)s * a.out * a.in;
      for (int kq = tid; kq < nkq; kq += HT) {
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int o = 0; o < OUT; ++o) {
          const float4 wv = ldcg4(wsr + (int64_t)o * a.in + 4 * kq);
          const float g = dyrow[o];
          d.x = fmaf(g, wv.x, d.x); d.y = fmaf(g, wv.y, d.y); d.z = fmaf(g, wv.z, d.z); d.w = fmaf(g, wv.w, d.w);
        }
        if (a.relu_in) {                         // gradient w.r.t. the PRE-activation of the layer below
          const float4 xv = ldcg4(a.x + (int64_t)r * a.in + 4 * kq);
          d.x = xv.x > 0.f ? d.x : 0.f; d.y = xv.y > 0.f ? d.y : 0.f; d.z = xv.z > 0.f ? d.z : 0.f; d.w = xv.w > 0.f ? d.w : 0.f;
        }
        *reinterpret_cast<float4 *>(a.dz_prev + (int64_t)r * a.in + 4 * kq) = d;
      }
    }
  }
  pdl_launch_dependents();

  // ---- part B: wgrad + the analytic epilogue for the CTA's columns ---------------------------------------------------------
  const int kq_lo = (int)((int64_t)blockIdx.x * nkq / gridDim.x), kq_hi = (int)((int64_t)(blockIdx.x + 1) * nkq / gridDim.x);
  for (int kq0 = kq_lo; kq0 < kq_hi; kq0 += KC / 4) {
    const int ncol = min(KC, 4 * (kq_hi - kq0)), k0 = 4 * kq0;
    float gm = 0.0f, gr = 0.0f;                   // thread (o, column): sums over the samples
    const int o_t = tid / KC, c_t = tid - o_t * KC;
    const bool owner = o_t < OUT && c_t < ncol;
    for (int s = 0; s < a.S; ++s) {
      // G_s[o][c] = sum_b dy[s][b][o] relu(x[s][b][k0 + c]), the batch in passes of HT rows
      float G = 0.0f;
      for (int b0 = 0; b0 < a.B; b0 += HT) {
        __syncthreads();
        const int b = b0 + tid;
        if (b < a.B) {
          const int64_t r = (int64_t)s * a.B + b;
#pragma unroll
          for (int o = 0; o < OUT; ++o) dy_s[tid][o] = __ldg(a.dy_in + r * OUT + o);
          for (int c = 0; c < ncol; c += 4) {
            float4 xv = ldcg4(a.x + r * a.in + k0 + c);
            if (a.relu_in) { xv.x = fmaxf(xv.x, 0.f); xv.y = fmaxf(xv.y, 0.f); xv.z = fmaxf(xv.z, 0.f); xv.w = fmaxf(xv.w, 0.f); }
            x_s[tid][c] = xv.x; x_s[tid][c + 1] = xv.y; x_s[tid][c + 2] = xv.z; x_s[tid][c + 3] = xv.w;
          }
        }
        __syncthreads();
        if (owner) {
          const int nb = min(HT, a.B - b0);
#pragma unroll 4
          for (int bb = 0; bb < nb; ++bb) G = fmaf(dy_s[bb][o_t], x_s[bb][c_t], G);
        }
      }
      if (owner) {
        const int64_t e = (int64_t)o_t * a.in + k0 + c_t;
        const float mu = __ldg(a.w_mu + e), rho = __ldg(a.w_rho + e);
        const float sg = softplus_f(rho);
        float ep = 0.0f;
        if (sample) {
          if (a.eps_w) ep = __ldg(a.eps_w + (int64_t)s * a.out * a.in + e);
          else ep = philox_normal1(rng, rng.tensor_w, rng.sample_base + (uint32_t)s, (uint64_t)e);
        }
        const float w = sample ? __fadd_rn(mu, __fmul_rn(sg, ep)) : mu;
        const float gps = a.gp * (a.gp_dev ? __ldg(a.gp_dev + s * a.g_dev_stride) : 1.0f);
        const float gqs = a.gq * (a.gq_dev ? __ldg(a.gq_dev + s * a.g_dev_stride) : 1.0f);
        float t = G;
        if (gps != 0.0f) t = fmaf(-gps * w, prior_R(a.prior, w), t);
        gm += t;
        gr += -expm1f(-sg) * (t * ep - gqs / sg);          // sigmoid(rho) = 1 - exp(-softplus(rho))
      }
    }
    if (owner) {
      const int64_t e = (int64_t)o_t * a.in + k0 + c_t;
      const bool acc = a.flags & BBB_F_ACCUM;
      a.g_w_mu[e] = acc ? fmaf(osc, gm, a.g_w_mu[e]) : osc * gm;
      a.g_w_rho[e] = acc ? fmaf(osc, gr, a.g_w_rho[e]) : osc * gr;
    }
  }
  // ---- bias gradients: CTA 0, thread o ----------------------------------------------------------------------------------------
  if (blockIdx.x == 0 && tid < OUT) {
    const float bmu = __ldg(a.b_mu + tid), bsg = softplus_f(__ldg(a.b_rho + tid));
    float gbm = 0.0f, gbr = 0.0f;
    for (int s = 0; s < a.S; ++s) {
      float col = 0.0f;
      for (int b = 0; b < a.B; ++b) col += __ldg(a.dy_in + ((int64_t)s * a.B + b) * OUT + tid);
      float ep = 0.0f;
      if (sample)
        ep = a.eps_b ? __ldg(a.eps_b + (int64_t)s * a.out + tid)
                     : philox_normal1(rng, rng.tensor_b, rng.sample_base + (uint32_t)s, (uint64_t)tid);
      const float bv = sample ? __fadd_rn(bmu, __fmul_rn(bsg, ep)) : bmu;
      const float gps = a.gp * (a.gp_dev ? __ldg(a.gp_dev + s * a.g_dev_stride) : 1.0f);
      const float gqs = a.gq * (a.gq_dev ? __ldg(a.gq_dev + s * a.g_dev_stride) : 1.0f);
      float t = col;
      if (gps != 0.0f) t = fmaf(-gps * bv, prior_R(a.prior, bv), t);
      gbm += t;
      gbr += -expm1f(-bsg) * (t * ep - gqs / bsg);
    }
    const bool acc = a.flags & BBB_F_ACCUM;
    a.g_b_mu[tid] = acc ? fmaf(osc, gbm, a.g_b_mu[tid]) : osc * gbm;
    a.g_b_rho[tid] = acc ? fmaf(osc, gbr, a.g_b_rho[tid]) : osc * gbr;
  }
}

template <int OUT>
int launch_fwd_t(const Head2Args &a, int grid, cudaStream_t st) {
  BBB_CHECK_CUDA(launch_pdl(head2_fwd_kernel<OUT>, dim3(grid), dim3(HT), 0, st, a));
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}
template <int OUT>
int launch_bwd_t(const Head2Args &a, int grid, cudaStream_t st) {
  BBB_CHECK_CUDA(launch_pdl(head2_bwd_kernel<OUT>, dim3(grid), dim3(HT), 0, st, a));
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}
#define HEAD2_DISPATCH(fn, a, grid, st)                                                      \
  switch ((a).out) {                                                                         \
    case 1: return fn<1>(a, grid, st);                                                       \
    case 2: return fn<2>(a, grid, st);                                                       \
    case 3: return fn<3>(a, grid, st);                                                       \
    case 4: return fn<4>(a, grid, st);                                                       \
    case 5: return fn<5>(a, grid, st);                                                       \
    case 6: return fn<6>(a, grid, st);                                                       \
    case 7: return fn<7>(a, grid, st);                                                       \
    case 8: return fn<8>(a, grid, st);                                                       \
    case 9: return fn<9>(a, grid, st);                                                       \
    case 10: return fn<10>(a, grid, st);                                                     \
    case 11: return fn<11>(a, grid, st);                                                     \
    case 12: return fn<12>(a, grid, st);                                                     \
    case 13: return fn<13>(a, grid, st);                                                     \
    case 14: return fn<14>(a, grid, st);                                                     \
    case 15: return fn<15>(a, grid, st);                                                     \
    default: return fn<16>(a, grid, st);                                                     \
  }

}  // namespace

bool head2_supported(int64_t S, int64_t B, int64_t in, int64_t out) {
  return out >= 1 && out <= MAXO && in >= 4 && in % 4 == 0 && in <= 8192 && S >= 1 && B >= 1 && S * B < (1 << 30);
}

// scratch floats the head needs: sampled weights + sampled biases
int64_t head2_scratch_floats(int64_t S, int64_t in, int64_t out) { return S * out * in + S * MAXO; }

int launch_head2_fwd(const MlpLayerDesc &l, int64_t S, int64_t B, const RngDev &rng, const PriorDev &prior, int flags,
                     int nll_kind, const void *target, float sigma, float grad_scale, float *d_out, double *logp,
                     double *logq, double *nll, float beta, const float *beta_dev, float *out4, uint32_t *done,
                     cudaStream_t st) {
  Head2Args a{};
  a.x = l.x; a.w_mu = l.w_mu; a.w_rho = l.w_rho; a.b_mu = l.b_mu; a.b_rho = l.b_rho; a.eps_w = l.eps_w; a.eps_b = l.eps_b;
  a.ws = l.w_sample; a.rng = rng; a.prior = prior;
  a.S = (int)S; a.B = (int)B; a.in = (int)l.in; a.out = (int)l.out; a.flags = flags; a.relu_in = (flags & BBB_F_RELU_IN) ? 1 : 0;
  a.nll_kind = nll_kind;
  if (nll_kind == BBB_NLL_CE) a.target_i = static_cast<const int64_t *>(target);
  if (nll_kind == BBB_NLL_GAUSS) {
    a.target_f = static_cast<const float *>(target);
    const double var = (double)sigma * sigma;
    a.inv_2var = (float)(0.5 / var); a.inv_var = (float)(1.0 / var);
    a.cst = (float)(log((double)sigma) + 0.918938533204672741780329736406);
  }
  a.grad_scale = grad_scale; a.y = l.y; a.dy = d_out; a.logp = logp; a.logq = logq; a.nll = nll;
  a.beta = beta; a.beta_dev = beta_dev; a.out4 = out4; a.done = done;
  int grid = (int)(S * B < sm_count() ? S * B : sm_count());       // one resident wave: the kernel contains a grid barrier
  HEAD2_DISPATCH(launch_fwd_t, a, grid, st);
}

int launch_head2_bwd(const MlpLayerDesc &l, int64_t S, int64_t B, const RngDev &rng, const PriorDev &prior, int flags,
                     float gp, float gq, const float *gp_dev, const float *gq_dev, int g_dev_stride,
                     const float *out_scale_dev, cudaStream_t st) {
  Head2Args a{};
  a.x = l.x; a.w_mu = l.w_mu; a.w_rho = l.w_rho; a.b_mu = l.b_mu; a.b_rho = l.b_rho; a.eps_w = l.eps_w; a.eps_b = l.eps_b;
  a.ws = l.w_sample; a.rng = rng; a.prior = prior;
  a.S = (int)S; a.B = (int)B; a.in = (int)l.in; a.out = (int)l.out; a.flags = flags; a.relu_in = (flags & BBB_F_RELU_IN) ? 1 : 0;
  a.dy_in = l.dz; a.gp = gp; a.gq = gq; a.gp_dev = gp_dev; a.gq_dev = gq_dev; a.g_dev_stride = g_dev_stride;
  a.out_scale_dev = out_scale_dev;
  a.dz_prev = l.dx; a.g_w_mu = l.g_w_mu; a.g_w_rho = l.g_w_rho; a.g_b_mu = l.g_b_mu; a.g_b_rho = l.g_b_rho;
  const int grid = sm_count();
  HEAD2_DISPATCH(launch_bwd_t, a, grid, st);
}

}  // namespace bbb
