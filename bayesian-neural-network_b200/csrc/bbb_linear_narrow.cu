// Weight-sampling Bayesian linear layer with a NARROW output (out <= 16): the classification / regression head
// (MNIST-shape: 1200 -> 10; regression and bandit nets: hidden -> 1).  Such a layer has a few thousand weights, so
// neither the tensor pipe nor HBM matters; what matters is that it does not cost a chain of latency-bound phases.
// Both kernels are exact fp32 (FMA contractions, precise softplus / log-densities) and are what the exact-fp32 mode
// runs for such layers (in TF32 mode the stream-K forward and the fused backward are faster: measured 11 / 19 us
// against 26 / 21 us at 1200 -> 10, because here every forward CTA samples the whole matrix).  No split-K, no
// atomics on the outputs.
//
//   forward   grid (row groups, S).  Every CTA samples the whole W_s = mu + sigma eps of its sample into shared
//             memory (in * out weights: ~2 us of work), then each warp takes batch rows and forms the <= 16 dot
//             products of a row with lanes striding over k (coalesced x, conflict-free W).  Row group 0 also
//             accumulates the log-prior / log-posterior of the sample.
//   backward  grid over 4-aligned ranges of input columns.  A CTA owns W[:, i_lo:i_hi) completely: per sample it
//             samples those weights once, and per 128-row batch chunk forms G[o][i] = sum_b dz[b][o] x[b][i] and
//             dX[b][i] = sum_o dz[b][o] W[o][i] (its columns of dx are complete: plain stores, optional (x > 0) mask),
//             then the analytic mu/rho-gradient epilogue.  CTA 0 also does the bias row.
#include "bbb_common.cuh"
#include "bbb_kernels.h"

namespace bbb {
namespace {

constexpr int NO = 16;    // widest output handled
constexpr int NTH = 256;  // threads per CTA
constexpr int BC = 128;   // batch chunk of the backward

struct WQuad {
  float mu[4], sg[4], ep[4], w[4];
};
// mu, sigma, eps, w of the 4 weights (o, i..i+3) of sample s; rows are 16-byte multiples (vec_in)
__device__ __forceinline__ void sample_wquad(const LinArgs &a, int s, int64_t e, bool sample, bool need_sigma, WQuad &q) {
  const float4 m = __ldg(reinterpret_cast<const float4 *>(a.w_mu + e));
  q.mu[0] = m.x; q.mu[1] = m.y; q.mu[2] = m.z; q.mu[3] = m.w;
  if (sample || need_sigma) {
    const float4 r = __ldg(reinterpret_cast<const float4 *>(a.w_rho + e));
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
__device__ __forceinline__ void bias_of(const LinArgs &a, int s, int64_t o, bool sample, bool need_sigma, float &b,
                                        float &sg, float &ep) {
  const float mu = __ldg(a.b_mu + o);
  sg = (sample || need_sigma) ? softplus_f(__ldg(a.b_rho + o)) : 0.0f;
  ep = 0.0f;
  if (sample)
    ep = a.eps_b ? __ldg(a.eps_b + (int64_t)s * a.out + o)
                 : philox_normal1(a.rng, a.rng.tensor_b, a.rng.sample_base + (uint32_t)s, (uint64_t)o);
  b = sample ? __fadd_rn(mu, __fmul_rn(sg, ep)) : mu;
}

// ==================================================================================================
// forward
// ==================================================================================================
template <bool kLogProb>
__global__ void __launch_bounds__(NTH) narrow_fwd_kernel(const LinArgs a_in, int rows_per_cta) {
  extern __shared__ __align__(16) float Ws[];  // [out][in]
  __shared__ float bias_s[NO];
  __shared__ float red[64];
  pdl_launch_dependents();
  pdl_wait();
  LinArgs a = a_in;
  rng_resolve(a.rng);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, s = blockIdx.y;
  const bool sample = a.flags & BBB_F_SAMPLE, relu = a.flags & BBB_F_RELU_IN;
  const bool lpcta = kLogProb && blockIdx.x == 0;
  const int in = (int)a.in, out = (int)a.out;
  float lp = 0.0f, lq = 0.0f;

  // 1. the sample's whole weight matrix, formed once per CTA
  for (int q4 = tid; q4 < (out * in) >> 2; q4 += NTH) {
    const int64_t e = (int64_t)q4 << 2;
    WQuad q;
    sample_wquad(a, s, e, sample, lpcta, q);
    *reinterpret_cast<float4 *>(Ws + e) = make_float4(q.w[0], q.w[1], q.w[2], q.w[3]);
    if (lpcta) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { lp += logp_elem(a.prior, q.w[j]); lq += logq_elem(q.sg[j], q.ep[j]); }
    }
  }
  if (tid < out) {
    float bv, sg, ep;
    bias_of(a, s, tid, sample, lpcta, bv, sg, ep);
    bias_s[tid] = bv;
    if (lpcta) { lp += logp_elem(a.prior, bv); lq += logq_elem(sg, ep); }
  }
  __syncthreads();

  // 2. one warp per batch row: lanes stride over k, <= 16 running dot products
  const int64_t r_end = min(a.B, (int64_t)(blockIdx.x + 1) * rows_per_cta);
  const float *xs = a.x + (int64_t)s * a.x_sstride;
  float *ys = a.y + (int64_t)s * a.B * a.out;
  for (int64_t b = (int64_t)blockIdx.x * rows_per_cta + warp; b < r_end; b += NTH / 32) {
    float acc[NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) acc[o] = 0.0f;
    const float *xr = xs + b * a.in;
    for (int k = lane * 4; k < in; k += 128) {
      float4 xv = __ldg(reinterpret_cast<const float4 *>(xr + k));
      if (relu) { xv.x = fmaxf(xv.x, 0.f); xv.y = fmaxf(xv.y, 0.f); xv.z = fmaxf(xv.z, 0.f); xv.w = fmaxf(xv.w, 0.f); }
#pragma unroll
      for (int o = 0; o < NO; ++o) {
        if (o < out) {
          const float4 wv = *reinterpret_cast<const float4 *>(Ws + o * in + k);
          acc[o] = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, acc[o]))));
        }
      }
    }
#pragma unroll
    for (int o = 0; o < NO; ++o) {
      if (o < out) {
        const float v = warp_sum(acc[o]);
        if (lane == o) ys[b * a.out + o] = v + bias_s[o];
      }
    }
  }
  if (lpcta) block_sum2_atomic(lp, lq, red, a.logp + s, a.logq + s);
}

// ==================================================================================================
// backward
// ==================================================================================================
__global__ void __launch_bounds__(NTH) narrow_bwd_kernel(const LinArgs a_in, int wq) {
  __shared__ __align__(16) float dz_s[BC][NO];       // dz chunk (masked)
  __shared__ __align__(16) float x_s[BC][32 + 1];    // x chunk, this CTA's columns (ReLU applied)
  __shared__ __align__(16) float Wt[NO][32];         // sampled weights of the CTA's columns, current sample
  __shared__ float G[NO][32];
  pdl_launch_dependents();
  pdl_wait();
  LinArgs a = a_in;
  rng_resolve(a.rng);
  const int tid = threadIdx.x;
  const bool sample = a.flags & BBB_F_SAMPLE, relu = a.flags & BBB_F_RELU_IN, wgrad = !(a.flags & BBB_F_NO_WGRAD);
  const bool want_dx = !(a.flags & BBB_F_NO_DX), dx_preact = a.flags & BBB_F_DX_PREACT, accum = a.flags & BBB_F_ACCUM;
  const int out = (int)a.out;
  const int64_t i_lo = (int64_t)blockIdx.x * wq * 4;
  const int w = (int)min((int64_t)wq * 4, a.in - i_lo), nq = w >> 2;   // columns / quads of this CTA
  const bool bias_cta = blockIdx.x == 0;
  const float osc = a.out_scale_dev ? __ldg(a.out_scale_dev) : 1.0f;
  const float dxs = ((a.flags & BBB_F_SCALE_DX) && a.out_scale_dev) ? osc : 1.0f;

  // quad thread t < out * nq owns the weights (o, i_lo + 4 iq .. +3) for the whole kernel
  const bool qthread = tid < out * nq;
  const int qo = qthread ? tid / nq : 0, qi = qthread ? tid - qo * nq : 0;
  const int64_t qe = (int64_t)qo * a.in + i_lo + qi * 4;
  float gm[4] = {0.f, 0.f, 0.f, 0.f}, gr[4] = {0.f, 0.f, 0.f, 0.f};
  float gbm = 0.0f, gbr = 0.0f;

  for (int s = 0; s < a.S; ++s) {
    const float gps = a.gp * (a.gp_dev ? __ldg(a.gp_dev + s * a.g_dev_stride) : 1.0f);
    const float gqs = a.gq * (a.gq_dev ? __ldg(a.gq_dev + s * a.g_dev_stride) : 1.0f);
    WQuad q;
    if (qthread) {
      sample_wquad(a, s, qe, sample, true, q);
      *reinterpret_cast<float4 *>(&Wt[qo][qi * 4]) = make_float4(q.w[0], q.w[1], q.w[2], q.w[3]);
    }
    for (int idx = tid; idx < NO * 32; idx += NTH) G[idx >> 5][idx & 31] = 0.0f;
    float colsum = 0.0f;
    const float *dys = a.dy + (int64_t)s * a.B * a.out, *mks = a.mask ? a.mask + (int64_t)s * a.B * a.out : nullptr;
    const float *xs = a.x + (int64_t)s * a.x_sstride;
    for (int64_t b0 = 0; b0 < a.B; b0 += BC) {
      const int nb = (int)min((int64_t)BC, a.B - b0);
      __syncthreads();  // previous chunk consumed; Wt / G of this sample written
      for (int idx = tid; idx < nb * out; idx += NTH) {
        const int b = idx / out, o = idx - b * out;
        float v = __ldg(dys + (b0 + b) * a.out + o);
        if (mks && !(__ldg(mks + (b0 + b) * a.out + o) > 0.0f)) v = 0.0f;
        dz_s[b][o] = v;
      }
      for (int idx = tid; idx < nb * nq; idx += NTH) {
        const int b = idx / nq, iq = idx - b * nq;
        float4 v = __ldg(reinterpret_cast<const float4 *>(xs + (b0 + b) * a.in + i_lo + iq * 4));
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        x_s[b][iq * 4 + 0] = v.x; x_s[b][iq * 4 + 1] = v.y; x_s[b][iq * 4 + 2] = v.z; x_s[b][iq * 4 + 3] = v.w;
      }
      __syncthreads();
      if (wgrad) {
        for (int idx = tid; idx < out * w; idx += NTH) {   // G[o][i] += sum_b dz[b][o] x[b][i]
          const int o = idx / w, i = idx - o * w;
          float g = 0.0f;
          for (int b = 0; b < nb; ++b) g = fmaf(dz_s[b][o], x_s[b][i], g);
          G[o][i] += g;
        }
        if (bias_cta && tid < out)
          for (int b = 0; b < nb; ++b) colsum += dz_s[b][tid];
      }
      if (want_dx) {                                         // dX[b][i] = sum_o dz[b][o] W[o][i]: complete
        float *dxr = a.dx + (int64_t)s * a.B * a.in + i_lo;
        for (int idx = tid; idx < nb * w; idx += NTH) {
          const int b = idx / w, i = idx - b * w;
          float d = 0.0f;
          for (int o = 0; o < out; ++o) d = fmaf(dz_s[b][o], Wt[o][i], d);
          d *= dxs;
          if (dx_preact && !(x_s[b][i] > 0.0f)) d = 0.0f;
          dxr[(b0 + b) * a.in + i] = d;
        }
      }
    }
    __syncthreads();
    if (wgrad && qthread) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float t = G[qo][qi * 4 + j];
        if (gps != 0.0f) t = fmaf(-gps * q.w[j], prior_R(a.prior, q.w[j]), t);
        gm[j] += t;
        gr[j] += -expm1f(-q.sg[j]) * (t * q.ep[j] - gqs / q.sg[j]);   // sigmoid(rho) = 1 - e^-sigma
      }
    }
    if (wgrad && bias_cta && tid < out) {
      float bv, sg, ep;
      bias_of(a, s, tid, sample, true, bv, sg, ep);
      float t = colsum;
      if (gps != 0.0f) t = fmaf(-gps * bv, prior_R(a.prior, bv), t);
      gbm += t;
      gbr += -expm1f(-sg) * (t * ep - gqs / sg);
    }
    __syncthreads();  // G and Wt are rewritten by the next sample
  }
  if (wgrad && qthread) {
    float4 *pm = reinterpret_cast<float4 *>(a.g_w_mu + qe), *pr = reinterpret_cast<float4 *>(a.g_w_rho + qe);
    float4 om = make_float4(0.f, 0.f, 0.f, 0.f), orr = om;
    if (accum) { om = *pm; orr = *pr; }
    *pm = make_float4(fmaf(osc, gm[0], om.x), fmaf(osc, gm[1], om.y), fmaf(osc, gm[2], om.z), fmaf(osc, gm[3], om.w));
    *pr = make_float4(fmaf(osc, gr[0], orr.x), fmaf(osc, gr[1], orr.y), fmaf(osc, gr[2], orr.z), fmaf(osc, gr[3], orr.w));
  }
  if (wgrad && bias_cta && tid < out) {
    a.g_b_mu[tid] = accum ? fmaf(osc, gbm, a.g_b_mu[tid]) : osc * gbm;
    a.g_b_rho[tid] = accum ? fmaf(osc, gbr, a.g_b_rho[tid]) : osc * gbr;
  }
}

// ==================================================================================================
// local-reparameterisation layer with a narrow output (weights [in, out], out <= 16), exact fp32
// ==================================================================================================
__device__ __forceinline__ float lr_eps_a(const LrArgs &a, int s, int64_t idx) {   // idx = b * out + o
  if (a.eps_a) return __ldg(a.eps_a + (int64_t)s * a.B * a.out + idx);
  if (a.vec_out) {                        // same coordinates as the quad-wise generators of the other LR kernels
    float e[4];
    philox_normal4(a.rng, a.rng.tensor_w, a.rng.sample_base + (uint32_t)s, (uint32_t)(idx >> 2), e);
    return e[idx & 3];
  }
  return philox_normal1(a.rng, a.rng.tensor_w, a.rng.sample_base + (uint32_t)s, (uint64_t)idx);
}
__device__ __forceinline__ float lr_eps_b(const LrArgs &a, int s, int64_t o) {
  return a.eps_b ? __ldg(a.eps_b + (int64_t)s * a.out + o)
                 : philox_normal1(a.rng, a.rng.tensor_b, a.rng.sample_base + (uint32_t)s, (uint64_t)o);
}
__device__ __forceinline__ float lr_kl_elem(float mu, float sg, float log_sp, float inv_sp2) {
  return 0.5f * (2.0f * (log_sp - logf(sg)) - 1.0f + (sg * sg + mu * mu) * inv_sp2);
}

// forward: grid (row groups, S); mu and sigma^2 of the whole layer in shared memory, one warp per batch row
constexpr int LNT = 1024;   // threads of the LR narrow forward: the layer's weights are staged once per CTA
__global__ void __launch_bounds__(LNT) lr_narrow_fwd_kernel(const LrArgs a_in, int rows_per_cta) {
  extern __shared__ __align__(16) float Wm[];   // [in][out] mu, then [in][out] sigma^2
  __shared__ float bias_s[NO];
  __shared__ float red[64];
  pdl_launch_dependents();
  pdl_wait();
  LrArgs a = a_in;
  rng_resolve(a.rng);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, s = blockIdx.y;
  const bool sample = a.flags & BBB_F_SAMPLE, relu = a.flags & BBB_F_RELU_IN;
  const bool klcta = (a.flags & BBB_F_LOGPROB) && blockIdx.x == 0 && s == 0;
  const int in = (int)a.in, out = (int)a.out, nw = in * out;
  float *Wv = Wm + nw;
  const float log_sp = logf(a.sigma_p), inv_sp2 = 1.0f / (a.sigma_p * a.sigma_p);
  float kl = 0.0f;
  for (int e = tid; e < nw; e += LNT) {
    const float mu = __ldg(a.w_mu + e);
    Wm[e] = mu;
    if (sample || klcta) {
      const float sg = softplus_f(__ldg(a.w_rho + e));
      Wv[e] = sg * sg;
      if (klcta) kl += lr_kl_elem(mu, sg, log_sp, inv_sp2);
    }
  }
  if (tid < out) {
    const float mu = __ldg(a.b_mu + tid);
    float bv = mu;
    if (sample || klcta) {
      const float sg = softplus_f(__ldg(a.b_rho + tid));
      if (sample) bv = fmaf(sg, lr_eps_b(a, s, tid), mu);
      if (klcta) kl += lr_kl_elem(mu, sg, log_sp, inv_sp2);
    }
    bias_s[tid] = bv;
  }
  __syncthreads();
  const int64_t r_end = min(a.B, (int64_t)(blockIdx.x + 1) * rows_per_cta);
  const float *xs = a.x + (int64_t)s * a.x_sstride;
  const int64_t base = (int64_t)s * a.B * a.out;
  for (int64_t b = (int64_t)blockIdx.x * rows_per_cta + warp; b < r_end; b += LNT / 32) {
    float g[NO], v[NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) g[o] = v[o] = 0.0f;
    for (int k0 = lane; k0 < in; k0 += 128) {
      float xq[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {                     // four independent loads in flight
        const int k = k0 + 32 * u;
        const float x = k < in ? __ldg(xs + b * a.in + k) : 0.0f;
        xq[u] = relu ? fmaxf(x, 0.0f) : x;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + 32 * u;
        if (k < in) {
          const float x = xq[u], x2 = x * x;
#pragma unroll
          for (int o = 0; o < NO; ++o) {
            if (o < out) {
              g[o] = fmaf(x, Wm[k * out + o], g[o]);
              if (sample) v[o] = fmaf(x2, Wv[k * out + o], v[o]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int o = 0; o < NO; ++o) {
      if (o < out) {
        const float gs = warp_sum(g[o]), vs = sample ? warp_sum(v[o]) : 0.0f;
        if (lane == o) {
          const float d = sample ? sqrtf(vs) : 0.0f;
          const float e = sample ? lr_eps_a(a, s, b * a.out + o) : 0.0f;
          a.y[base + b * a.out + o] = gs + d * e + bias_s[o];
          if (a.delta) a.delta[base + b * a.out + o] = d;
        }
      }
    }
  }
  if (klcta) block_sum2_atomic(kl, 0.0f, red, a.kl, nullptr);
}

// backward: grid over ranges of <= 32 weight rows (input features); a CTA owns mu/rho[i_lo:i_hi, :] completely
__global__ void __launch_bounds__(NTH) lr_narrow_bwd_kernel(const LrArgs a_in, int wrows) {
  __shared__ float dz_s[BC][NO], dv_s[BC][NO];
  __shared__ float x_s[BC][32 + 1];
  __shared__ float Mu[32][NO], S2[32][NO];   // this CTA's rows of mu and sigma^2
  pdl_launch_dependents();
  pdl_wait();
  LrArgs a = a_in;
  rng_resolve(a.rng);
  const int tid = threadIdx.x;
  const bool sample = a.flags & BBB_F_SAMPLE, relu = a.flags & BBB_F_RELU_IN, wgrad = !(a.flags & BBB_F_NO_WGRAD);
  const bool want_dx = !(a.flags & BBB_F_NO_DX), dx_preact = a.flags & BBB_F_DX_PREACT, accum = a.flags & BBB_F_ACCUM;
  const int out = (int)a.out;
  const int64_t i_lo = (int64_t)blockIdx.x * wrows;
  const int w = (int)min((int64_t)wrows, a.in - i_lo);
  const bool bias_cta = blockIdx.x == 0;
  const float osc = a.out_scale_dev ? __ldg(a.out_scale_dev) : 1.0f;
  const float dxs = ((a.flags & BBB_F_SCALE_DX) && a.out_scale_dev) ? osc : 1.0f;
  const float gk = (a.flags & BBB_F_LOGPROB) ? a.g_kl * (a.g_kl_dev ? __ldg(a.g_kl_dev) : 1.0f) : 0.0f;
  const float inv_sp2 = 1.0f / (a.sigma_p * a.sigma_p);
  // weight thread t < w * out owns element (i_lo + t / out, t % out)
  const bool wthread = tid < w * out;
  const int wi = wthread ? tid / out : 0, wo = wthread ? tid - wi * out : 0;
  const int64_t we = (i_lo + wi) * a.out + wo;
  float mu = 0.0f, sg = 1.0f;
  if (wthread) {
    mu = __ldg(a.w_mu + we);
    sg = softplus_f(__ldg(a.w_rho + we));
    Mu[wi][wo] = mu;
    S2[wi][wo] = sg * sg;
  }
  float g1 = 0.0f, g2 = 0.0f, gbmu = 0.0f, gbrho = 0.0f;
  for (int s = 0; s < a.S; ++s) {
    const int64_t base = (int64_t)s * a.B * a.out;
    const float *xs = a.x + (int64_t)s * a.x_sstride;
    float colsum = 0.0f;
    for (int64_t b0 = 0; b0 < a.B; b0 += BC) {
      const int nb = (int)min((int64_t)BC, a.B - b0);
      __syncthreads();
      for (int idx = tid; idx < nb * out; idx += NTH) {
        const int b = idx / out, o = idx - b * out;
        const int64_t e = (b0 + b) * a.out + o;
        float z = __ldg(a.dy + base + e);
        if (a.mask && !(__ldg(a.mask + base + e) > 0.0f)) z = 0.0f;
        float dv = 0.0f;
        if (sample) {
          const float d = __ldg(a.delta_in + base + e);
          if (d > 0.0f) dv = z * lr_eps_a(a, s, e) / (2.0f * d);
        }
        dz_s[b][o] = z;
        dv_s[b][o] = dv;
      }
      for (int idx = tid; idx < nb * w; idx += NTH) {
        const int b = idx / w, i = idx - b * w;
        float x = __ldg(xs + (b0 + b) * a.in + i_lo + i);
        x_s[b][i] = relu ? fmaxf(x, 0.0f) : x;
      }
      __syncthreads();
      if (wgrad && wthread) {
        for (int b = 0; b < nb; ++b) {
          const float x = x_s[b][wi];
          g1 = fmaf(x, dz_s[b][wo], g1);
          g2 = fmaf(x * x, dv_s[b][wo], g2);
        }
      }
      if (wgrad && bias_cta && tid < out)
        for (int b = 0; b < nb; ++b) colsum += dz_s[b][tid];
      if (want_dx) {
        float *dxr = a.dx + (int64_t)s * a.B * a.in + i_lo;
        for (int idx = tid; idx < nb * w; idx += NTH) {
          const int b = idx / w, i = idx - b * w;
          float p1 = 0.0f, p2 = 0.0f;
          for (int o = 0; o < out; ++o) {
            p1 = fmaf(dz_s[b][o], Mu[i][o], p1);
            p2 = fmaf(dv_s[b][o], S2[i][o], p2);
          }
          const float x = x_s[b][i];
          float d = dxs * fmaf(2.0f * x, p2, p1);
          if (dx_preact && !(x > 0.0f)) d = 0.0f;
          dxr[(b0 + b) * a.in + i] = d;
        }
      }
    }
    if (wgrad && bias_cta && tid < out) {
      gbmu += colsum;
      if (sample) gbrho += colsum * lr_eps_b(a, s, tid);
    }
  }
  if (wgrad && wthread) {
    const float gm = fmaf(gk * mu, inv_sp2, g1);
    const float gr = -expm1f(-sg) * (2.0f * sg * g2 + gk * (sg * inv_sp2 - 1.0f / sg));
    a.g_w_mu[we] = accum ? fmaf(osc, gm, a.g_w_mu[we]) : osc * gm;
    a.g_w_rho[we] = accum ? fmaf(osc, gr, a.g_w_rho[we]) : osc * gr;
  }
  if (wgrad && bias_cta && tid < out) {
    const float bmu = a.b_mu[tid], bsg = softplus_f(a.b_rho[tid]);
    const float gm = fmaf(gk * bmu, inv_sp2, gbmu);
    const float gr = -expm1f(-bsg) * (gbrho + gk * (bsg * inv_sp2 - 1.0f / bsg));
    a.g_b_mu[tid] = accum ? fmaf(osc, gm, a.g_b_mu[tid]) : osc * gm;
    a.g_b_rho[tid] = accum ? fmaf(osc, gr, a.g_b_rho[tid]) : osc * gr;
  }
}

inline int cdiv_i(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
constexpr int kMaxWsBytes = 200 * 1024;

}  // namespace

bool linear_narrow_supported(const LinArgs &a) {
  return a.out >= 1 && a.out <= NO && a.vec_in && a.in >= 4 && a.B >= 1 && a.S >= 1 &&
         a.in * a.out * (int64_t)sizeof(float) <= kMaxWsBytes;
}

int launch_linear_fwd_narrow(const LinArgs &a, cudaStream_t st) {
  // about two CTAs per SM over (row groups x samples), at least 8 rows (one per warp) each
  int rows = cdiv_i(a.B * a.S, 2 * sm_count());
  rows = ((rows < 8 ? 8 : rows) + 7) / 8 * 8;
  dim3 grid(cdiv_i(a.B, rows), (unsigned)a.S);
  const size_t smem = (size_t)a.in * a.out * sizeof(float);
  if (a.flags & BBB_F_LOGPROB) {
    BBB_CHECK_CUDA(cudaFuncSetAttribute(narrow_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxWsBytes));
    BBB_CHECK_CUDA(launch_pdl(narrow_fwd_kernel<true>, grid, dim3(NTH), smem, st, a, rows));
  } else {
    BBB_CHECK_CUDA(cudaFuncSetAttribute(narrow_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxWsBytes));
    BBB_CHECK_CUDA(launch_pdl(narrow_fwd_kernel<false>, grid, dim3(NTH), smem, st, a, rows));
  }
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}

int launch_linear_bwd_narrow(const LinArgs &a, cudaStream_t st) {
  const int nq_i = (int)(a.in / 4);
  int wq = cdiv_i(nq_i, sm_count());           // about one column range per SM, at most 32 columns
  if (wq > 8) wq = 8;
  if (wq < 1) wq = 1;
  BBB_CHECK_CUDA(launch_pdl(narrow_bwd_kernel, dim3(cdiv_i(nq_i, wq)), dim3(NTH), 0, st, a, wq));
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}

bool lr_narrow_supported(const LrArgs &a) {
  return a.out >= 1 && a.out <= NO && a.in >= 1 && a.B >= 1 && a.S >= 1 &&
         2 * a.in * a.out * (int64_t)sizeof(float) <= kMaxWsBytes;
}

int launch_lr_fwd_narrow(const LrArgs &a, cudaStream_t st) {
  int rows = cdiv_i(a.B * a.S, 2 * sm_count());
  rows = ((rows < 64 ? 64 : rows) + 31) / 32 * 32;      // >= 2 rows per warp: the weight staging is per CTA
  dim3 grid(cdiv_i(a.B, rows), (unsigned)a.S);
  const size_t smem = 2 * (size_t)a.in * a.out * sizeof(float);
  BBB_CHECK_CUDA(cudaFuncSetAttribute(lr_narrow_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxWsBytes));
  BBB_CHECK_CUDA(launch_pdl(lr_narrow_fwd_kernel, grid, dim3(LNT), smem, st, a, rows));
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}

int launch_lr_bwd_narrow(const LrArgs &a, cudaStream_t st) {
  int wrows = cdiv_i(a.in, sm_count());        // about one range of weight rows per SM
  const int cap = NTH / (int)a.out < 32 ? NTH / (int)a.out : 32;   // one thread per owned weight
  if (wrows > cap) wrows = cap;
  if (wrows < 1) wrows = 1;
  BBB_CHECK_CUDA(launch_pdl(lr_narrow_bwd_kernel, dim3(cdiv_i(a.in, wrows)), dim3(NTH), 0, st, a, wrows));
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}

}  // namespace bbb
