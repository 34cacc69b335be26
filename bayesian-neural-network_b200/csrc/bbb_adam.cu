// Multi-tensor Adam: torch.optim.Adam.step() for every parameter tensor of the network in ONE launch
// (SURVEY 8f-1).  The stock optimiser issues several foreach kernels per step; at the MNIST-shape config
// those passes cost more than the whole fused forward+backward.  HBM-bound: 16 B read + 12 B written per
// parameter (p, g, m, v in; p, m, v out), 16-byte vector accesses, grid = a multiple of the 148 SMs.
// The update rule is exactly torch's (non-amsgrad, no weight decay):
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
#include "bbb_common.cuh"

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

__global__ void __launch_bounds__(256) adam_kernel(const AdamArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ AdamConst cs;
  if (threadIdx.x == 0) cs = adam_consts(a.lr, a.b1, a.b2, a.eps, a.step, a.step_dev, a.lr_scale_dev);
  __syncthreads();
  const AdamConst c = cs;
  const int64_t total = a.qend[a.n - 1], stride = (int64_t)gridDim.x * blockDim.x;
  int ti = 0;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += stride) {
    while (q >= a.qend[ti]) ++ti;
    const int64_t e = (q - (ti ? a.qend[ti - 1] : 0)) * 4;
    const int valid = (int)min((int64_t)4, a.size[ti] - e);
    float *pp = a.p[ti] + e, *pm = a.m[ti] + e, *pv = a.v[ti] + e;
    const float *pg = a.g[ti] + e;
    if (valid == 4 && a.vec[ti]) {
      float4 P = *reinterpret_cast<float4 *>(pp), M = *reinterpret_cast<float4 *>(pm), V = *reinterpret_cast<float4 *>(pv);
      const float4 G = *reinterpret_cast<const float4 *>(pg);
      adam1(P.x, G.x, M.x, V.x, c);
      adam1(P.y, G.y, M.y, V.y, c);
      adam1(P.z, G.z, M.z, V.z, c);
      adam1(P.w, G.w, M.w, V.w, c);
      *reinterpret_cast<float4 *>(pp) = P;
      *reinterpret_cast<float4 *>(pm) = M;
      *reinterpret_cast<float4 *>(pv) = V;
    } else {
      for (int j = 0; j < valid; ++j) {
        float P = pp[j], M = pm[j], V = pv[j];
        adam1(P, pg[j], M, V, c);
        pp[j] = P; pm[j] = M; pv[j] = V;
      }
    }
  }
}

}  // namespace
}  // namespace bbb

using namespace bbb;

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
  const int64_t cap = (int64_t)kSMs * 8;
  if (blocks > cap) blocks = cap;
  BBB_CHECK_CUDA(launch_pdl(adam_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, a));
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}
