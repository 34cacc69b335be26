// Fused backward of the weight-sampling Bayesian linear layer for batches of at most 128 rows (tcgen05 kind::tf32):
// ONE kernel per layer does wgrad, the analytic mu/rho-gradient epilogue and dgrad, and regenerates eps once.
//
// A CTA owns a block of the weight matrix: rows [o_t0, o_t0 + rows) of one output-row tile (<= 128 rows) and a
// 4-aligned column range [i_lo, i_hi) of at most 88 columns.  Per Monte-Carlo sample it
//   1. stages the activation tiles in shared memory in their NATURAL row-major orientation, as stacked
//      [128 batch rows][128 B] swizzled regions:  dz_s[b][o_t0 .. +128)  and  x_s[b][i_lo .. +96);
//   2. MMA1  G[o][i] = sum_b dz[b][o] x[b][i]    -- both operands MN-major (the batch is K), accumulator in TMEM;
//   3. drains G to shared memory in the layout of an MN-major operand tile  Wt[o][i];
//   4. epilogue, one weight quad per thread-iteration, coalesced along i: regenerate eps (Philox), w = mu + sigma eps,
//      t = G - gp w R(w);  grad_mu += t;  grad_rho += sigmoid(rho) (t eps - gq / sigma)   (sums kept in registers
//      across samples), and overwrite the G quad IN PLACE by the TF32 weights w -- the tile is now the B operand
//      of dgrad, zero outside the CTA's own block;
//   5. MMA2  dX[b][i] = sum_o dz[b][o] Wt[o][i]  -- A = the same dz tile read K-major, B = Wt MN-major;
//   6. drains dX, optionally multiplied by (x > 0) (the ReLU that produced this layer's input), and adds the
//      partial sum over its o rows into dx with red.global.add.v4.f32.
// No transposes anywhere: a [rows = batch][32 floats] region is an MN-major operand as it stands.  MN-major TF32
// operands use the SWIZZLE_128B_BASE32B format and K-major ones SWIZZLE_128B (bbb_tc.cuh), so the dz tile is stored
// twice, once per swizzle, from the same registers.  eps is generated once per weight and sample for the whole
// backward, and W never leaves the SM.
//
// Work split: the output rows are cut into equal tiles of T_o <= 128 rows, each tile's input columns are divided
// evenly (granularity one quad) over grid.x CTAs; grid.x * grid.y is about one CTA per SM (launch_bwd_fused).
#include "bbb_tc_tiles.cuh"

namespace bbb {
namespace {

using namespace tc;
using namespace tcx;

constexpr int FT = 512;              // threads per CTA (16 warps, one CTA per SM)
constexpr int REG = 128 * 128;       // bytes of one [128 rows][128 B] swizzled region
constexpr int DZ_REGS = 4;           // dz tile: 128 o columns
constexpr int XW_REGS = 3;           // x tile / Wt tile: 96 i columns
constexpr int NWIN = XW_REGS * 32;   // MMA N
constexpr int WMAX = 88;             // widest column range a CTA may own
constexpr int QMAX = 6;              // weight quads per thread: 512 * 6 * 4 >= 128 * 88
constexpr int kFusedDyn = (2 * DZ_REGS + 2 * XW_REGS) * REG + 1024;  // dz (MN) | dz (K) | x | Wt
constexpr uint32_t kFusedTmemCols = 256;  // G: columns [0, 96), dX: columns [128, 224)
static_assert(FT * QMAX * 4 >= 128 * WMAX, "every own weight needs a thread slot");

struct FCtl {
  uint64_t bar;
  uint32_t tmem_base;
};

// address of float4 column `col4` of row `row` in an MN-major (BASE32B) / a K-major (SWIZZLE_128B) tile
__device__ __forceinline__ uint8_t *mn_ptr(uint8_t *tile, int row, int col4) {
  return tile + (col4 >> 3) * REG + mn32_off(row, col4 & 7);
}
__device__ __forceinline__ uint8_t *k_ptr(uint8_t *tile, int row, int col4) {
  return tile + (col4 >> 3) * REG + sw128_off(row, col4 & 7);
}

template <bool kDx>
__global__ void __launch_bounds__(FT, 1) bwd_fused_kernel(const LinArgs a_in, int T_o) {
  extern __shared__ uint8_t dsm[];
  __shared__ FCtl ctl;
  __shared__ float csum_part[2][128];   // (the dynamic tiles leave 2 KB of the 227 KB)
  LinArgs a = a_in;
  rng_resolve(a.rng);
  uint8_t *dz_t = align1024(dsm), *dzk_t = dz_t + DZ_REGS * REG, *x_t = dzk_t + DZ_REGS * REG, *w_t = x_t + XW_REGS * REG;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool sample = a.flags & BBB_F_SAMPLE, relu = a.flags & BBB_F_RELU_IN, wgrad = !(a.flags & BBB_F_NO_WGRAD);
  const bool dx_preact = a.flags & BBB_F_DX_PREACT, accum = a.flags & BBB_F_ACCUM;
  const int B = (int)a.B;

  // this CTA's block
  const int64_t o_t0 = (int64_t)blockIdx.y * T_o;
  const int rows = (int)min((int64_t)T_o, a.out - o_t0);
  const int nq_i = (int)(a.in >> 2);
  const int q_lo = (int)((int64_t)blockIdx.x * nq_i / gridDim.x), q_hi = (int)((int64_t)(blockIdx.x + 1) * nq_i / gridDim.x);
  const int tq = q_hi - q_lo;                      // own quads per row (<= 22)
  const int64_t i_lo = (int64_t)q_lo * 4;
  const int nquad = rows * tq;
  const bool bias_cta = blockIdx.x == 0;

  if (warp == 0) tmem_alloc(smem_u32(&ctl.tmem_base), kFusedTmemCols);
  if (tid == 32) {
    mbar_init(smem_u32(&ctl.bar), 1);
    mbar_fence_init();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = ctl.tmem_base, tmem_g = tmem, tmem_dx = tmem + 128;
  constexpr uint32_t idesc1 = idesc_tf32_major(128, NWIN, 1, 1);  // MMA1: A = dz (MN-major), B = x (MN-major)
  constexpr uint32_t idesc2 = idesc_tf32_major(128, NWIN, 0, 1);  // MMA2: A = dz (K-major),  B = Wt (MN-major)
  const float osc = a.out_scale_dev ? __ldg(a.out_scale_dev) : 1.0f;
  const float dxs = ((a.flags & BBB_F_SCALE_DX) && a.out_scale_dev) ? osc : 1.0f;

  float gm[QMAX][4], gr[QMAX][4];
#pragma unroll
  for (int j = 0; j < QMAX; ++j)
#pragma unroll
    for (int c = 0; c < 4; ++c) gm[j][c] = gr[j][c] = 0.0f;
  float gbm = 0.0f, gbr = 0.0f;
  uint32_t phase = 0;

  for (int s = 0; s < a.S; ++s) {
    // ---- 1. stage dz_s [128 b][128 o] and x_s [128 b][96 i] (loads first, then the swizzled stores) ----------
    {
      const int64_t zbase = (int64_t)s * a.B * a.out;
      const int ncol4 = (rows + 3) >> 2;             // dz columns beyond the tile's rows are never used
      float4 zv[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int idx = tid + FT * r, b = idx >> 5, c4 = idx & 31;
        const int64_t o = o_t0 + c4 * 4;
        zv[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.vec_out) {
          zv[r] = ld_row4_al(a.dy + zbase, c4 < ncol4 ? b : a.B, o, a.B, a.out);
          if (a.mask) zv[r] = mask4(zv[r], ld_row4_al(a.mask + zbase, c4 < ncol4 ? b : a.B, o, a.B, a.out));
        } else if (c4 < ncol4) {
          zv[r] = ld_row4(a.dy + zbase, b, o, a.B, a.out, false);
          if (a.mask) zv[r] = mask4(zv[r], ld_row4(a.mask + zbase, b, o, a.B, a.out, false));
        }
      }
      const bool load_x = s == 0 || a.x_sstride != 0;
      float4 xv[6];
      if (load_x) {
        const float *xs = a.x + (int64_t)s * a.x_sstride;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          const int idx = tid + FT * r, b = idx / 24, c4 = idx - b * 24;
          xv[r] = ld_row4_al(xs, c4 < tq ? b : a.B, i_lo + c4 * 4, a.B, a.in);   // columns beyond the own range: zero
        }
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int idx = tid + FT * r;
        const float4 v = zv[r];
        *reinterpret_cast<float4 *>(mn_ptr(dz_t, idx >> 5, idx & 31)) = v;
        if (kDx) *reinterpret_cast<float4 *>(k_ptr(dzk_t, idx >> 5, idx & 31)) = v;
      }
      if (load_x) {
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          const int idx = tid + FT * r, b = idx / 24;
          *reinterpret_cast<float4 *>(mn_ptr(x_t, b, idx - b * 24)) = relu ? relu4(xv[r]) : xv[r];
        }
      }
    }
    fence_proxy_async_smem();
    __syncthreads();

    // ---- 2. MMA1: G[o][i] = sum_b dz[b][o] x[b][i], K = batch in steps of 8 rows -------------------------------
    if (tid == 0) {
      tc_fence_after_sync();
      const int nk = (B + 7) >> 3;
      for (int k8 = 0; k8 < nk; ++k8)
        mma_tf32(tmem_g, smem_desc_mn32(smem_u32(dz_t) + k8 * 1024, REG), smem_desc_mn32(smem_u32(x_t) + k8 * 1024, REG),
                 idesc1, k8 > 0 ? 1u : 0u);
      mma_commit(smem_u32(&ctl.bar));
    }
    // the first weight quad's parameters travel while the tensor pipe works
    float4 nmu = make_float4(0.f, 0.f, 0.f, 0.f), nrho = nmu;
    if (tid < nquad) {
      const int r = tid / tq;
      const int64_t e = (o_t0 + r) * a.in + i_lo + (int64_t)(tid - r * tq) * 4;
      nmu = __ldg(reinterpret_cast<const float4 *>(a.w_mu + e));
      nrho = __ldg(reinterpret_cast<const float4 *>(a.w_rho + e));
    }
    mbar_wait(smem_u32(&ctl.bar), phase);
    phase ^= 1u;
    tc_fence_after_sync();

    // ---- 3. G: TMEM [lane = o][column = i] -> Wt layout in shared memory ------------------------------------------
    {
      const int o_r = (warp & 3) * 32 + lane, cg = warp >> 2;   // 4 column groups of 24
#pragma unroll
      for (int c0 = 0; c0 < 24; c0 += 8) {
        const int col = cg * 24 + c0;
        float v[8];
        tmem_ld8(tmem_g + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)col, v);
        *reinterpret_cast<float4 *>(mn_ptr(w_t, o_r, col >> 2)) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4 *>(mn_ptr(w_t, o_r, (col >> 2) + 1)) = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
    tc_fence_before_sync();
    __syncthreads();

    // ---- 4. epilogue over the CTA's own quads; G quad -> w quad in place ------------------------------------------
    const float gps = a.gp * (a.gp_dev ? __ldg(a.gp_dev + s * a.g_dev_stride) : 1.0f);
    const float gqs = a.gq * (a.gq_dev ? __ldg(a.gq_dev + s * a.g_dev_stride) : 1.0f);
#pragma unroll
    for (int j = 0; j < QMAX; ++j) {
      const int q = tid + FT * j;
      if (q < nquad) {
        const float4 cmu = nmu, crho = nrho;
        const int r = q / tq, iq = q - r * tq;
        const int64_t e = (o_t0 + r) * a.in + i_lo + (int64_t)iq * 4;
        if (j + 1 < QMAX && q + FT < nquad) {
          const int rn = (q + FT) / tq;
          const int64_t en = (o_t0 + rn) * a.in + i_lo + (int64_t)(q + FT - rn * tq) * 4;
          nmu = __ldg(reinterpret_cast<const float4 *>(a.w_mu + en));
          nrho = __ldg(reinterpret_cast<const float4 *>(a.w_rho + en));
        }
        Quad qd;
        qd.mu[0] = cmu.x; qd.mu[1] = cmu.y; qd.mu[2] = cmu.z; qd.mu[3] = cmu.w;
        qd.rho[0] = crho.x; qd.rho[1] = crho.y; qd.rho[2] = crho.z; qd.rho[3] = crho.w;
        float sgm[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) softplus_sigmoid_fast(qd.rho[c], qd.sg[c], sgm[c]);
        float ep[4], w[4];
        sample_quad(a, s, e, qd, sample, ep, w);
        float4 *slot = reinterpret_cast<float4 *>(mn_ptr(w_t, r, iq));
        if (wgrad) {
          const float4 G = *slot;
          const float Gv[4] = {G.x, G.y, G.z, G.w};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float t = Gv[c];
            if (gps != 0.0f) t = fmaf(-gps * w[c], prior_R_fast(a.prior, w[c]), t);
            gm[j][c] += t;
            gr[j][c] += sgm[c] * fmaf(t, ep[c], -gqs * __fdividef(1.0f, qd.sg[c]));
          }
        }
        if (kDx) *slot = make_float4(to_tf32(w[0]), to_tf32(w[1]), to_tf32(w[2]), to_tf32(w[3]));
      }
    }
    if (kDx) {  // everything in the 128 x 96 window that is not an own weight must not contribute to dX
      for (int idx = tid; idx < 128 * (NWIN / 4); idx += FT) {
        const int r = idx / (NWIN / 4), c4 = idx - r * (NWIN / 4);
        if (r >= rows || c4 >= tq) *reinterpret_cast<float4 *>(mn_ptr(w_t, r, c4)) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    // bias gradients, part 1: column sums of the dz tile, 2 row groups x 128 columns over the first 256 threads
    if (wgrad && bias_cta && tid < 256) {
      const int o_l = tid & 127, grp = tid >> 7;
      const uint8_t *zp = dz_t + (o_l >> 5) * REG + (o_l & 3) * 4;
      const int b_end = min(B, grp * 64 + 64);
      float part = 0.0f;
#pragma unroll 8
      for (int b = grp * 64; b < b_end; ++b) part += *reinterpret_cast<const float *>(zp + mn32_off(b, (o_l >> 2) & 7));
      csum_part[grp][o_l] = part;
    }
    // part 2 (after the next CTA-wide barrier, overlapping MMA2): prior/posterior terms of the bias row
    auto bias_finish = [&]() {
      if (wgrad && bias_cta && tid < rows) {
        const float colsum = csum_part[0][tid] + csum_part[1][tid];
        const int64_t o = o_t0 + tid;
        float bv, sg, ep;
        bias_elem(a, s, o, sample, true, bv, sg, ep);
        float t = colsum;
        if (gps != 0.0f) t = fmaf(-gps * bv, prior_R(a.prior, bv), t);
        gbm += t;
        gbr += -expm1f(-sg) * (t * ep - gqs / sg);
      }
    };

    if (kDx) {
      fence_proxy_async_smem();
      __syncthreads();
      // ---- 5. MMA2: dX[b][i] = sum_o dz[b][o] Wt[o][i], K = o in steps of 8 rows of Wt ---------------------------
      if (tid == 0) {
        tc_fence_after_sync();
        const int nk = (rows + 7) >> 3;
        for (int k8 = 0; k8 < nk; ++k8)
          mma_tf32(tmem_dx, smem_desc_sw128(smem_u32(dzk_t) + (k8 >> 2) * REG + (k8 & 3) * 32),
                   smem_desc_mn32(smem_u32(w_t) + k8 * 1024, REG), idesc2, k8 > 0 ? 1u : 0u);
        mma_commit(smem_u32(&ctl.bar));
      }
      bias_finish();
      mbar_wait(smem_u32(&ctl.bar), phase);
      phase ^= 1u;
      tc_fence_after_sync();
      // ---- 6. dX: TMEM [lane = b][column = i] -> (x > 0) mask -> red.add into dx --------------------------------
      {
        const int b = (warp & 3) * 32 + lane, cg = warp >> 2;
        float *drow = a.dx + ((int64_t)s * a.B + b) * a.in + i_lo;
#pragma unroll
        for (int c0 = 0; c0 < 24; c0 += 8) {
          const int col = cg * 24 + c0;
          float v[8];
          tmem_ld8(tmem_dx + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)col, v);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c4 = (col >> 2) + h;
            if (b < B && c4 < tq) {
              float4 d = make_float4(dxs * v[4 * h], dxs * v[4 * h + 1], dxs * v[4 * h + 2], dxs * v[4 * h + 3]);
              if (dx_preact) d = mask4(d, *reinterpret_cast<const float4 *>(mn_ptr(x_t, b, c4)));
              red_add_v4(drow + c4 * 4, d);
            }
          }
        }
      }
      tc_fence_before_sync();
    } else if (wgrad && bias_cta) {
      __syncthreads();
      bias_finish();
    }
    __syncthreads();  // the tiles and the accumulators are reused by the next sample
  }

  // ---- parameter gradients ------------------------------------------------------------------------------------
  if (wgrad) {
#pragma unroll
    for (int j = 0; j < QMAX; ++j) {
      const int q = tid + FT * j;
      if (q < nquad) {
        const int r = q / tq, iq = q - r * tq;
        const int64_t e = (o_t0 + r) * a.in + i_lo + (int64_t)iq * 4;
        float4 *pm = reinterpret_cast<float4 *>(a.g_w_mu + e), *pr = reinterpret_cast<float4 *>(a.g_w_rho + e);
        float4 om = make_float4(0.f, 0.f, 0.f, 0.f), orr = om;
        if (accum) { om = *pm; orr = *pr; }
        *pm = make_float4(fmaf(osc, gm[j][0], om.x), fmaf(osc, gm[j][1], om.y), fmaf(osc, gm[j][2], om.z), fmaf(osc, gm[j][3], om.w));
        *pr = make_float4(fmaf(osc, gr[j][0], orr.x), fmaf(osc, gr[j][1], orr.y), fmaf(osc, gr[j][2], orr.z), fmaf(osc, gr[j][3], orr.w));
      }
    }
    if (bias_cta && tid < rows) {
      const int64_t o = o_t0 + tid;
      a.g_b_mu[o] = accum ? fmaf(osc, gbm, a.g_b_mu[o]) : osc * gbm;
      a.g_b_rho[o] = accum ? fmaf(osc, gbr, a.g_b_rho[o]) : osc * gbr;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kFusedTmemCols);
}

inline int cdiv_i(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace

bool linear_bwd_fused_supported(const LinArgs &a) {
  return a.vec_in && a.in >= 4 && a.out >= 1 && a.B >= 1 && a.B <= 128 && a.S >= 1;
}

int launch_linear_bwd_fused(const LinArgs &a, cudaStream_t st) {
  const bool want_dx = !(a.flags & BBB_F_NO_DX);
  const int n_ot = cdiv_i(a.out, 128);
  const int T_o = ((cdiv_i(a.out, n_ot) + 3) / 4) * 4;         // equal row tiles; a multiple of 4 keeps dz loads vectorised
  int n_c = kSMs / n_ot;                                       // about one CTA per SM
  const int need = cdiv_i(a.in / 4, WMAX / 4);                 // every column range must fit the 88-column limit
  if (n_c < need) n_c = need;
  if (n_c > a.in / 4) n_c = (int)(a.in / 4);
  if (n_c < 1) n_c = 1;
  dim3 grid(n_c, cdiv_i(a.out, T_o));
  if (want_dx && !(a.flags & BBB_F_OUT_ZEROED)) {
    BBB_CHECK_CUDA(cudaMemsetAsync(a.dx, 0, sizeof(float) * (size_t)a.S * a.B * a.in, st));
    note_launch();
  }
  if (want_dx) {
    BBB_CHECK_CUDA(cudaFuncSetAttribute(bwd_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedDyn));
    bwd_fused_kernel<true><<<grid, FT, kFusedDyn, st>>>(a, T_o);
  } else {
    BBB_CHECK_CUDA(cudaFuncSetAttribute(bwd_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedDyn));
    bwd_fused_kernel<false><<<grid, FT, kFusedDyn, st>>>(a, T_o);
  }
  BBB_CHECK_LAUNCH();
  return BBB_OK;
}

}  // namespace bbb
