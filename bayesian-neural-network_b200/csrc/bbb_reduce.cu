// Bandwidth-bound single-pass kernels: ELBO reductions over (mu, rho[, eps]), closed-form KL,
// the Philox eps stream, likelihood terms and ELBO assembly.
//
// Roofline: HBM.  bbb_logprob_reduce reads 8 B/weight (mu, rho; +4 with injected eps, +4 when W is
// written) and does ~60 flops + Philox per weight; bbb_kl_gauss reads 8 B/weight.  Loads are 16-byte
// vectors, four independent vectors in flight per thread, grid = a multiple of the 148 SMs.
#include <stdarg.h>

#include <atomic>

#include "bbb_common.cuh"

namespace bbb {

// ---- error plumbing ------------------------------------------------------------------
char *last_error_buf() {
  static thread_local char buf[512] = "ok";
  return buf;
}
int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

static std::atomic<uint64_t> g_launches{0};
void note_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

namespace {

constexpr int RT = 256;  // threads per CTA of the reduction kernels

template <bool kInjected, bool kWriteW>
__global__ void __launch_bounds__(RT) logprob_reduce_kernel(const float *__restrict__ mu,
                                                            const float *__restrict__ rho,
                                                            const float *__restrict__ eps, RngDev rng,
                                                            uint32_t tensor, uint32_t sample, PriorDev prior,
                                                            int64_t n, int sample_flag, float *__restrict__ w_out,
                                                            double *logp, double *logq) {
  __shared__ float red[64];
  rng_resolve(rng);
  float lp = 0.0f, lq = 0.0f;
  const int64_t nq = n >> 2;  // whole quads
  const int64_t stride = (int64_t)gridDim.x * RT;
  for (int64_t q = (int64_t)blockIdx.x * RT + threadIdx.x; q < nq; q += stride) {
    const float4 m4 = __ldg(reinterpret_cast<const float4 *>(mu) + q);
    const float4 r4 = __ldg(reinterpret_cast<const float4 *>(rho) + q);
    const float m[4] = {m4.x, m4.y, m4.z, m4.w}, r[4] = {r4.x, r4.y, r4.z, r4.w};
    float e[4] = {0.f, 0.f, 0.f, 0.f}, w[4];
    if (sample_flag) {
      if (kInjected) {
        const float4 e4 = __ldg(reinterpret_cast<const float4 *>(eps) + q);
        e[0] = e4.x; e[1] = e4.y; e[2] = e4.z; e[3] = e4.w;
      } else {
        philox_normal4(rng, tensor, sample, (uint32_t)q, e);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float sg = softplus_f(r[j]);
      w[j] = __fadd_rn(m[j], __fmul_rn(sg, e[j]));
      lp += logp_elem(prior, w[j]);
      lq += logq_elem(sg, e[j]);
    }
    if (kWriteW) reinterpret_cast<float4 *>(w_out)[q] = make_float4(w[0], w[1], w[2], w[3]);
  }
  // tail (n % 4 elements) handled by the first threads of block 0
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (nq << 2) + threadIdx.x;
    const float sg = softplus_f(rho[i]);
    float e = 0.0f;
    if (sample_flag) e = kInjected ? eps[i] : philox_normal1(rng, tensor, sample, (uint64_t)i);
    const float w = __fadd_rn(mu[i], __fmul_rn(sg, e));
    lp += logp_elem(prior, w);
    lq += logq_elem(sg, e);
    if (kWriteW) w_out[i] = w;
  }
  block_sum2_atomic(lp, lq, red, logp, logq);
}

__global__ void __launch_bounds__(RT) kl_gauss_kernel(const float *__restrict__ mu, const float *__restrict__ rho,
                                                      float log_sp, float inv_sp2, int64_t n, double *kl) {
  __shared__ float red[64];
  float acc = 0.0f;
  const int64_t nq = n >> 2, stride = (int64_t)gridDim.x * RT;
  for (int64_t q = (int64_t)blockIdx.x * RT + threadIdx.x; q < nq; q += stride) {
    const float4 m4 = __ldg(reinterpret_cast<const float4 *>(mu) + q);
    const float4 r4 = __ldg(reinterpret_cast<const float4 *>(rho) + q);
    const float m[4] = {m4.x, m4.y, m4.z, m4.w}, r[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float sg = softplus_f(r[j]);
      acc += 2.0f * (log_sp - logf(sg)) - 1.0f + (sg * sg + m[j] * m[j]) * inv_sp2;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (nq << 2) + threadIdx.x;
    const float sg = softplus_f(rho[i]);
    acc += 2.0f * (log_sp - logf(sg)) - 1.0f + (sg * sg + mu[i] * mu[i]) * inv_sp2;
  }
  block_sum2_atomic(0.5f * acc, 0.0f, red, kl, nullptr);
}

__global__ void __launch_bounds__(RT) philox_fill_kernel(float *__restrict__ out, int64_t n, RngDev rng,
                                                         uint32_t tensor, uint32_t sample) {
  const int64_t nq = (n + 3) >> 2, stride = (int64_t)gridDim.x * RT;
  for (int64_t q = (int64_t)blockIdx.x * RT + threadIdx.x; q < nq; q += stride) {
    float z[4];
    philox_normal4(rng, tensor, sample, (uint32_t)q, z);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if ((q << 2) + j < n) out[(q << 2) + j] = z[j];
  }
}

// one warp per (s, b) row of logits
__global__ void __launch_bounds__(128) nll_ce_kernel(const float *__restrict__ logits,
                                                     const int64_t *__restrict__ target, int64_t rows, int64_t B,
                                                     int64_t C, float grad_scale, double *nll,
                                                     float *__restrict__ dlogits) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float local = 0.0f;
  for (int64_t row = (int64_t)blockIdx.x * 4 + warp; row < rows; row += (int64_t)gridDim.x * 4) {
    const float *z = logits + row * C;
    const int64_t t = target[row % B];
    float mx = -INFINITY;
    for (int64_t c = lane; c < C; c += 32) mx = fmaxf(mx, z[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.0f;
    for (int64_t c = lane; c < C; c += 32) se += expf(z[c] - mx);
    se = warp_sum(se);
    const float lse = mx + logf(se);
    if (lane == 0) local += lse - z[t];
    if (dlogits) {
      float *g = dlogits + row * C;
      for (int64_t c = lane; c < C; c += 32) g[c] = grad_scale * (expf(z[c] - lse) - (c == t ? 1.0f : 0.0f));
    }
  }
  block_sum2_atomic(local, 0.0f, red, nll, nullptr);
}

__global__ void __launch_bounds__(RT) nll_gauss_kernel(const float *__restrict__ out, const float *__restrict__ target,
                                                       float inv_2var, float inv_var, float cst, int64_t n,
                                                       int64_t per_sample, float grad_scale, double *nll,
                                                       float *__restrict__ dout) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[64];
  float acc = 0.0f;
  for (int64_t i = (int64_t)blockIdx.x * RT + threadIdx.x; i < n; i += (int64_t)gridDim.x * RT) {
    const float d = out[i] - target[i % per_sample];
    acc += fmaf(d * d, inv_2var, cst);
    if (dout) dout[i] = grad_scale * d * inv_var;
  }
  block_sum2_atomic(acc, 0.0f, red, nll, nullptr);
}

__global__ void elbo_finalize_kernel(const double *logp, const double *logq, const double *kl, const double *nll,
                                     int S, float beta, const float *beta_dev, float *out4) {
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (beta_dev) beta *= *beta_dev;
  const float nll_m = (float)(nll[0] / S);
  if (kl) {
    const float k = (float)kl[0];
    out4[0] = beta * k + nll_m; out4[1] = k; out4[2] = nll_m; out4[3] = 0.0f;
  } else {
    double lp = 0.0, lq = 0.0;
    for (int s = 0; s < S; ++s) { lp += (double)(float)logp[s]; lq += (double)(float)logq[s]; }
    const float lpm = (float)(lp / S), lqm = (float)(lq / S);
    out4[0] = beta * lqm - beta * lpm + nll_m; out4[1] = lpm; out4[2] = lqm; out4[3] = nll_m;
  }
}

// signal-to-noise ratio in decibels of every Gaussian parameter, 10 log10(|mu| / softplus(rho)) (weight_pruning.py:81-83),
// and/or the pruning pass of weight_pruning.py:85-115: mu, rho *= (snr > threshold), in place.  One pass over (mu, rho).
__global__ void __launch_bounds__(RT) snr_kernel(float *__restrict__ mu, float *__restrict__ rho, int64_t n,
                                                 float *__restrict__ snr_out, int prune, float threshold,
                                                 unsigned long long *kept) {
  const int64_t stride = (int64_t)gridDim.x * RT;
  unsigned long long mine = 0;
  for (int64_t i = (int64_t)blockIdx.x * RT + threadIdx.x; i < n; i += stride) {
    const float m = mu[i], r = rho[i];
    const float snr = 10.0f * log10f(fabsf(m) / softplus_f(r));
    if (snr_out) snr_out[i] = snr;
    if (prune) {
      const bool keep = snr > threshold;        // (NaN and -inf, i.e. mu == 0, are dropped, as `snrs > threshold` does)
      if (!keep) { mu[i] = m * 0.0f; rho[i] = r * 0.0f; }
      mine += keep ? 1ull : 0ull;
    }
  }
  if (prune && kept) {
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(kept, mine);
  }
}

// BNN_Classification.predict (class_task.py:81-87): softmax of every sampled forward, averaged over the samples.
// One warp per batch row; logits [S,B,C], probs [B,C].
__global__ void __launch_bounds__(128) softmax_mean_kernel(const float *__restrict__ logits, int64_t S, int64_t B, int64_t C,
                                                           float *__restrict__ probs) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t b = (int64_t)blockIdx.x * 4 + warp; b < B; b += (int64_t)gridDim.x * 4) {
    for (int64_t c0 = 0; c0 < C; c0 += 32) {        // (C <= 32: one pass; wider heads accumulate chunk by chunk)
      const int64_t c = c0 + lane;
      float acc = 0.0f;
      for (int64_t s = 0; s < S; ++s) {
        const float *z = logits + (s * B + b) * C;
        float mx = -INFINITY;
        for (int64_t j = lane; j < C; j += 32) mx = fmaxf(mx, z[j]);
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.0f;
        for (int64_t j = lane; j < C; j += 32) sum += expf(z[j] - mx);
        sum = warp_sum(sum);
        if (c < C) acc += expf(z[c] - mx) / sum;
      }
      if (c < C) probs[b * C + c] = acc / (float)S;
    }
  }
}

inline int grid_for(int64_t work_items, int per_cta) {
  int64_t need = (work_items + per_cta - 1) / per_cta;
  int64_t cap = (int64_t)sm_count() * 8;  // 8 resident CTAs of 256 threads per SM
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}
inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace
}  // namespace bbb

using namespace bbb;

extern "C" int bbb_logprob_reduce(const float *mu, const float *rho, const float *eps, uint64_t seed, uint32_t step,
                                  uint32_t sample, uint32_t tensor_id, const bbb_prior *prior, int64_t n,
                                  int32_t flags, float *w_out, double *logp, double *logq, void *stream) {
  BBB_CHECK_ARG(mu && rho && prior && logp && logq, "null pointer");
  BBB_CHECK_ARG(n >= 0, "negative size");
  BBB_CHECK_ARG(aligned16(mu) && aligned16(rho) && (!eps || aligned16(eps)) && (!w_out || aligned16(w_out)),
                "pointers must be 16-byte aligned");
  if (n == 0) return BBB_OK;
  bbb_rng r{seed, step, 0, 0, nullptr};
  RngDev rng = make_rng_dev(&r);
  PriorDev pd = make_prior_dev(prior);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(n >> 2, RT * 4);
  const int sf = (flags & BBB_F_SAMPLE) ? 1 : 0;
#define LAUNCH(INJ, WR) \
  logprob_reduce_kernel<INJ, WR><<<grid, RT, 0, st>>>(mu, rho, eps, rng, tensor_id, sample, pd, n, sf, w_out, logp, logq)
  if (eps && w_out) LAUNCH(true, true);
  else if (eps) LAUNCH(true, false);
  else if (w_out) LAUNCH(false, true);
  else LAUNCH(false, false);
#undef LAUNCH
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}

extern "C" int bbb_kl_gauss(const float *mu, const float *rho, float sigma_p, int64_t n, double *kl, void *stream) {
  BBB_CHECK_ARG(mu && rho && kl, "null pointer");
  BBB_CHECK_ARG(n >= 0 && sigma_p > 0, "bad size or sigma_p");
  BBB_CHECK_ARG(aligned16(mu) && aligned16(rho), "pointers must be 16-byte aligned");
  if (n == 0) return BBB_OK;
  kl_gauss_kernel<<<grid_for(n >> 2, RT * 4), RT, 0, (cudaStream_t)stream>>>(
      mu, rho, logf(sigma_p), 1.0f / (sigma_p * sigma_p), n, kl);
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}

extern "C" int bbb_philox_fill_normal(float *out, int64_t n, uint64_t seed, uint32_t step, uint32_t sample,
                                      uint32_t tensor_id, void *stream) {
  BBB_CHECK_ARG(out && n >= 0, "null pointer or negative size");
  if (n == 0) return BBB_OK;
  bbb_rng r{seed, step, 0, 0, nullptr};
  philox_fill_kernel<<<grid_for((n + 3) >> 2, RT * 4), RT, 0, (cudaStream_t)stream>>>(out, n, make_rng_dev(&r),
                                                                                    tensor_id, sample);
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}

extern "C" int bbb_nll_ce(const float *logits, const int64_t *target, int64_t S, int64_t B, int64_t C,
                          float grad_scale, double *nll, float *dlogits, void *stream) {
  BBB_CHECK_ARG(logits && target && nll, "null pointer");
  BBB_CHECK_ARG(S >= 0 && B >= 0 && C > 0, "bad shape");
  const int64_t rows = S * B;
  if (rows == 0) return BBB_OK;
  BBB_CHECK_CUDA(launch_pdl(nll_ce_kernel, dim3(grid_for(rows, 4)), dim3(128), 0, (cudaStream_t)stream, logits, target, rows, B,
                            C, grad_scale, nll, dlogits));
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}

extern "C" int bbb_nll_gauss(const float *out, const float *target, float sigma, int64_t S, int64_t B, int64_t D,
                             float grad_scale, double *nll, float *dout, void *stream) {
  BBB_CHECK_ARG(out && target && nll, "null pointer");
  BBB_CHECK_ARG(S >= 0 && B >= 0 && D >= 0 && sigma > 0, "bad shape or sigma");
  const int64_t n = S * B * D;
  if (n == 0) return BBB_OK;
  const double var = (double)sigma * sigma;
  BBB_CHECK_CUDA(launch_pdl(nll_gauss_kernel, dim3(grid_for(n, RT)), dim3(RT), 0, (cudaStream_t)stream, out, target,
                            (float)(0.5 / var), (float)(1.0 / var),
                            (float)(log((double)sigma) + 0.918938533204672741780329736406), n, B * D, grad_scale, nll, dout));
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}

extern "C" int bbb_elbo_finalize(const double *logp, const double *logq, const double *kl, const double *nll,
                                 int64_t S, float beta, const float *beta_dev, float *out4, void *stream) {
  BBB_CHECK_ARG(nll && out4 && S > 0, "null pointer or S <= 0");
  BBB_CHECK_ARG(kl || (logp && logq), "need kl or logp+logq");
  BBB_CHECK_CUDA(launch_pdl(elbo_finalize_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, logp, logq, kl, nll, (int)S, beta,
                            beta_dev, out4));
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}

namespace bbb { namespace {
__global__ void counter_add_kernel(uint32_t *c, uint32_t inc) {
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0 && blockIdx.x == 0) *c += inc;
}
} }
extern "C" int bbb_counter_add(uint32_t *counter, uint32_t inc, void *stream) {
  BBB_CHECK_ARG(counter, "null pointer");
  BBB_CHECK_CUDA(launch_pdl(counter_add_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, counter, inc));
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}

extern "C" uint64_t bbb_launch_count(void) { return bbb::g_launches.load(std::memory_order_relaxed); }
extern "C" int bbb_version(void) { return BBB_VERSION; }
extern "C" const char *bbb_last_error_string(void) { return last_error_buf(); }

// ---- consumers of the (mu, rho) stream and of the sampled outputs ---------------------------------
extern "C" int bbb_snr(const float *mu, const float *rho, int64_t n, float *snr_out, void *stream) {
  BBB_CHECK_ARG(n >= 0 && ((mu && rho && snr_out) || n == 0), "null pointer or negative size");
  if (n == 0) return BBB_OK;
  snr_kernel<<<grid_for(n, RT * 4), RT, 0, (cudaStream_t)stream>>>(const_cast<float *>(mu), const_cast<float *>(rho), n,
                                                                    snr_out, 0, 0.0f, nullptr);
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}

extern "C" int bbb_snr_prune(float *mu, float *rho, int64_t n, float threshold_db, unsigned long long *kept, void *stream) {
  BBB_CHECK_ARG(n >= 0 && ((mu && rho) || n == 0), "null pointer or negative size");
  if (n == 0) return BBB_OK;
  snr_kernel<<<grid_for(n, RT * 4), RT, 0, (cudaStream_t)stream>>>(mu, rho, n, nullptr, 1, threshold_db, kept);
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}

extern "C" int bbb_softmax_mean(const float *logits, int64_t S, int64_t B, int64_t C, float *probs, void *stream) {
  BBB_CHECK_ARG(S >= 1 && B >= 0 && C >= 1 && ((logits && probs) || B == 0), "null pointer or bad shape");
  if (B == 0) return BBB_OK;
  softmax_mean_kernel<<<grid_for(B, 4), 128, 0, (cudaStream_t)stream>>>(logits, S, B, C, probs);
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}
