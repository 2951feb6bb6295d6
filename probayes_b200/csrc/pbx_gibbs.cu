// pbx_gibbs.cu -- K5: batched Gibbs over the conditionals of a multivariate normal
// and the batched mvn density.
//
// Gibbs step (probayes/rf.py:446-462 -> rf_utils.py:50-65 -> cond_cov.py:42-65),
// one coordinate i = step mod d per step (tsteps=1):
//     u   = cdf_lo_i + (cdf_hi_i - cdf_lo_i) * r          r ~ U(0,1)  (vtypes.py:186)
//     x_i = ndtri(u) * stdv_i + mean_i + coef_i . (x_-i - mean_-i)
// with the truncation limits fixed at construction (cond_cov.py:38-39).  The dot
// product is evaluated as c0_i + coef_i . x with c0_i = mean_i - coef_i . mean
// folded on the host (coef_ii = 0).
//
// Kernel: four lanes = one chain, the state vector lives in REGISTERS (DP/4 doubles
// per lane, coordinate loop fully unrolled), the coefficient rows come from shared
// memory with 128-bit loads, the dot product finishes with two shuffle-adds.
//
// Density (probayes/prob.py:347-360 -> scipy multivariate_normal): maha = |(x -
// mean) W|^2.  For d >= 64 the [chains, d] x [d, d] whitening product runs on the
// FP64 tensor cores (mma.sync.m8n8k4.f64 -- tcgen05 has no FP64 kind), 8 chains per
// MMA row block; smaller d uses a plain FMA kernel.
#include <math.h>
#include "pbx_common.cuh"

#ifndef GB_THREADS
#define GB_THREADS 256                 // measured: 0.137 / 0.116 / 0.108 ms per sweep at 64 / 128 / 256
#endif

struct GibbsArgs {
  int C, d, T, thin;
  int64_t step0, chain0;
  uint64_t seed;
  const double* coef;     // [d][d]
  const double* c0;       // [d] (workspace: mean_i - coef_i . mean)
  const double* stdv;
  const double* cdf_lo;
  const double* cdf_hi;
  double* state;          // [d][C]
  const double* inj_runif;
  double* out_x;
  double* stat_sum;
  double* stat_sumsq;
};

__global__ void gibbs_c0_kernel(const double* coef, const double* mean, int d, double* c0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d) return;
  double acc = 0.0;
  for (int j = 0; j < d; ++j) acc = fma(coef[(int64_t)i * d + j], mean[j], acc);
  c0[i] = mean[i] - acc;
}

// State-independent part of one coordinate update, kept out of line: the coordinate
// loop is fully unrolled (register-resident state) and inlining ~250 instructions of
// Philox + normcdfinv at every site made the kernel 60 k instructions long and
// instruction-fetch bound (ncu: stall_no_inst 34 %).
__device__ __noinline__ double gibbs_draw_z(uint64_t seed, uint64_t gk, uint32_t gchain,
                                            const double* inj, double lo, double w, double sd,
                                            double c0) {
  double r;
  if (inj) {
    r = *inj;
  } else {
    pbx_u4 b = pbx_block(seed, gk, gchain, 0u);
    r = pbx_u52(b.x, b.y);
  }
  return normcdfinv(lo + w * r) * sd + c0;
}

// One chain = FOUR lanes (a warp = 8 chains): lane q of a chain keeps coordinates
// j = 4m + q in registers (DQ = DP/4 doubles instead of DP, so the kernel fits ~64
// registers instead of 255 + local-memory spills), the dot product coef_i . x is DQ
// FMAs per lane + two shuffle-adds, and the expensive state-independent part of a
// coordinate update -- Philox, ndtri(u) * stdv -- is computed by each lane for ITS OWN
// next coordinate, i.e. four coordinates' worth in parallel with no redundancy.
// kSweepRec: records fall on sweep boundaries only (thin % d == 0, call aligned to
// sweeps) -> one copy of the record code per sweep instead of one per coordinate.
// NT threads per CTA (256, or 128 when there are too few chains to give every SM a CTA of
// 64 chains: strong-scaled shards)
template <int DQ, bool kSweepRec, int NT = GB_THREADS>
__global__ void __launch_bounds__(NT)
    gibbs_mvn_kernel(const GibbsArgs a) {
  constexpr int DP = 4 * DQ;
  constexpr int RS = DQ + 2;                // padded row stride: the 4 sub-lane rows of a
                                            // coordinate land in distinct shared-memory banks
  extern __shared__ __align__(16) double sm[];
  double* s_coef = sm;                      // [DP][4][RS]: s_coef[i][q][m] = coef[i][4m + q]
  double* s_c0 = sm + DP * 4 * RS;          // [DP]
  double* s_sd = s_c0 + DP;
  double* s_lo = s_sd + DP;
  double* s_w = s_lo + DP;                  // hi - lo
  const int d = a.d;
  for (int idx = threadIdx.x; idx < DP * 4 * RS; idx += NT) {
    const int i = idx / (4 * RS), r = idx % (4 * RS), q = r / RS, mm = r % RS;
    const int j = 4 * mm + q;
    s_coef[idx] = (i < d && mm < DQ && j < d) ? a.coef[(int64_t)i * d + j] : 0.0;
  }
  for (int i = threadIdx.x; i < DP; i += NT) {
    const bool v = i < d;
    s_c0[i] = v ? a.c0[i] : 0.0;
    s_sd[i] = v ? a.stdv[i] : 0.0;
    s_lo[i] = v ? a.cdf_lo[i] : 0.0;
    s_w[i] = v ? a.cdf_hi[i] - a.cdf_lo[i] : 0.0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, q = lane & 3;
  const int64_t C = a.C;
  const int64_t c_raw = ((int64_t)blockIdx.x * (NT / 32) + (threadIdx.x >> 5)) * 8 + (lane >> 2);
  const bool valid = c_raw < C;
  const int64_t c = valid ? c_raw : C - 1;          // clamp: idle lanes stay in the shuffles
  const uint32_t gchain = (uint32_t)(a.chain0 + c);

  double x[DQ];
#pragma unroll
  for (int mm = 0; mm < DQ; ++mm) {
    const int j = 4 * mm + q;
    x[mm] = (j < d) ? a.state[(int64_t)j * C + c] : 0.0;
  }
  const bool stats = a.stat_sum != nullptr;
  const int64_t k_begin = a.step0, k_end = a.step0 + a.T;
  int until_rec = a.thin;
  int64_t rec = 0;
  auto record = [&]() {
    if (valid) {
#pragma unroll
      for (int mm = 0; mm < DQ; ++mm) {
        const int j = 4 * mm + q;
        if (j < d) {
          if (a.out_x) a.out_x[(rec * d + j) * C + c] = x[mm];
          if (stats) {                              // running sums over RECORDED states
            a.stat_sum[(int64_t)j * C + c] += x[mm];
            a.stat_sumsq[(int64_t)j * C + c] =
                fma(x[mm], x[mm], a.stat_sumsq[(int64_t)j * C + c]);
          }
        }
      }
    }
    ++rec;
  };
  for (int64_t sweep = k_begin / d; sweep * d < k_end; ++sweep) {
#pragma unroll
    for (int l = 0; l < DQ; ++l) {
      // ---- my own next coordinate i = 4l + q: draw and transform (state independent)
      double z = 0.0;
      {
        const int i = 4 * l + q;
        const int64_t gk = sweep * d + i;
        if (i < d && gk >= k_begin && gk < k_end)
          z = gibbs_draw_z(a.seed, (uint64_t)gk, gchain,
                           a.inj_runif ? a.inj_runif + (gk - k_begin) * C + c : nullptr,
                           s_lo[i], s_w[i], s_sd[i], s_c0[i]);
      }
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        const int i = 4 * l + o;
        const int64_t gk = sweep * d + i;
        if (i < d && gk >= k_begin && gk < k_end) {        // warp-uniform
          // conditional mean: coef_i . x  (my DQ coordinates, then the 4 sub-lanes)
          const double2* row = reinterpret_cast<const double2*>(s_coef + (i * 4 + q) * RS);
          double p0 = 0.0, p1 = 0.0;
          if (DQ >= 2) {
#pragma unroll
            for (int mm = 0; mm < DQ; mm += 2) {
              const double2 cf = row[mm / 2];
              p0 = fma(cf.x, x[mm], p0);
              p1 = fma(cf.y, x[mm + 1], p1);
            }
          } else {
            p0 = s_coef[(i * 4 + q) * RS] * x[0];
          }
          double dot = p0 + p1;
          dot += __shfl_xor_sync(0xffffffffu, dot, 1);
          dot += __shfl_xor_sync(0xffffffffu, dot, 2);
          if (q == o) x[l] = z + dot;                      // owner lane: x_i = ndtri*sd + c0 + dot
          if (!kSweepRec && --until_rec == 0) {
            until_rec = a.thin;
            record();
          }
        }
      }
    }
    if (kSweepRec) {
      until_rec -= d;
      if (until_rec == 0) {
        until_rec = a.thin;
        record();
      }
    }
  }
  if (valid) {
#pragma unroll
    for (int mm = 0; mm < DQ; ++mm) {
      const int j = 4 * mm + q;
      if (j < d) a.state[(int64_t)j * C + c] = x[mm];
    }
  }
}

// ---------------------------------------------------------------------------
// batched mvn density, plain FMA version: one thread per point, W from smem
// ---------------------------------------------------------------------------
// points: R blocks of [d][C]; point (r, c) -> x[(r*d + j)*C + c], out[r*C + c]
__global__ void __launch_bounds__(128)
    mvn_logpdf_kernel(const double* __restrict__ x, int d, int64_t C, int64_t R,
                      const double* __restrict__ mean, const double* __restrict__ W, double norm_c,
                      int log_pscale, double* __restrict__ out) {
  extern __shared__ __align__(16) double sm[];
  double* s_W = sm;            // [d][d]
  double* s_m = sm + d * d;    // [d]
  for (int i = threadIdx.x; i < d * d; i += blockDim.x) s_W[i] = W[i];
  for (int i = threadIdx.x; i < d; i += blockDim.x) s_m[i] = mean[i];
  __syncthreads();
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= C * R) return;
  const int64_t r = p / C, c = p % C;
  const double* xb = x + r * d * C + c;
  double maha = 0.0;
  for (int k = 0; k < d; ++k) {
    double y = 0.0;
    for (int j = 0; j < d; ++j) y = fma(xb[(int64_t)j * C] - s_m[j], s_W[j * d + k], y);
    maha = fma(y, y, maha);
  }
  const double lp = -0.5 * (norm_c + maha);
  out[p] = log_pscale ? lp : exp(lp);
}

// ---------------------------------------------------------------------------
// batched mvn density on the FP64 tensor cores, d = 64.
// One warp = 8 points per pass.  D[8 x 8] += A[8 x 4] * B[4 x 8] with
//   A[r][k] = x[k][c0 + r] - mean[k]   (thread T: r = T/4, k = T%4)
//   B[k][n] = W[k][n]                  (thread T: k = T%4, n = T/4)
//   D[r][n]: thread T holds (r = T/4, n = 2*(T%4) + {0,1})
// ---------------------------------------------------------------------------
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(128)
    mvn_logpdf_mma64_kernel(const double* __restrict__ x, int64_t C, int64_t R,
                            const double* __restrict__ mean, const double* __restrict__ W,
                            double norm_c, int log_pscale, double* __restrict__ out) {
  constexpr int D = 64;
  // W in shared memory, row stride 68 doubles (= 4 mod 16 double-banks): the B
  // fragment load (k = 4ks + q, n = 8nb + r) is bank-conflict free per half-warp
  __shared__ double s_W[D][D + 4];
  __shared__ double s_m[D];
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) s_W[i / D][i % D] = W[i];
  for (int i = threadIdx.x; i < D; i += blockDim.x) s_m[i] = mean[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = lane >> 2, q = lane & 3;
  const int64_t bpr = (C + 7) / 8, n_blocks = bpr * R;       // 8-point blocks per record
  for (int64_t blk = (int64_t)blockIdx.x * 4 + warp; blk < n_blocks; blk += (int64_t)gridDim.x * 4) {
    const int64_t rr = blk / bpr;
    const int64_t c = (blk % bpr) * 8 + r;
    const bool valid = c < C;
    const double* xb = x + rr * D * C;
    // this thread's A elements for the 16 k-steps: k = 4*ks + q
    double av[16];
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) {
      const int k = 4 * ks + q;
      av[ks] = valid ? xb[(int64_t)k * C + c] - s_m[k] : 0.0;
    }
    double maha = 0.0;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      double d0 = 0.0, d1 = 0.0;
#pragma unroll
      for (int ks = 0; ks < 16; ++ks) dmma_m8n8k4(d0, d1, av[ks], s_W[4 * ks + q][nb * 8 + r]);
      maha = fma(d0, d0, maha);
      maha = fma(d1, d1, maha);
    }
    // sum the 4 threads of a row (they hold different output columns)
    maha += __shfl_xor_sync(0xffffffffu, maha, 1);
    maha += __shfl_xor_sync(0xffffffffu, maha, 2);
    if (q == 0 && valid) {
      const double lp = -0.5 * (norm_c + maha);
      out[rr * C + c] = log_pscale ? lp : exp(lp);
    }
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static int mvn_logpdf_launch(pbx_ctx* ctx, const double* x, int d, int64_t C, int64_t R,
                             const double* mean, const double* W, double norm_c, int log_pscale,
                             double* out) {
  if (C * R == 0) return PBX_OK;
  if (d == 64) {
    int64_t want = ((C + 7) / 8 * R + 3) / 4;
    int grid = (int)(want > (int64_t)ctx->sm_count * 6 ? (int64_t)ctx->sm_count * 6 : want);
    mvn_logpdf_mma64_kernel<<<grid, 128, 0, ctx->stream>>>(x, C, R, mean, W, norm_c, log_pscale,
                                                          out);
  } else {
    const size_t smem = ((size_t)d * d + d) * sizeof(double);
    const int64_t grid = (C * R + 127) / 128;
    if (smem > 48 * 1024)
      PBX_CUDA(cudaFuncSetAttribute(mvn_logpdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
    mvn_logpdf_kernel<<<(unsigned)grid, 128, smem, ctx->stream>>>(x, d, C, R, mean, W, norm_c,
                                                                 log_pscale, out);
  }
  PBX_LAUNCH_CHECK(ctx);
  return PBX_OK;
}

extern "C" int pbx_mvn_logpdf(pbx_ctx* ctx, const double* x, int32_t n_dims, int64_t n_chains,
                              const double* mean, const double* whiten, double norm_c,
                              int32_t log_pscale, double* out) {
  PBX_REQUIRE(ctx && x && mean && whiten && out, "pbx_mvn_logpdf: null argument");
  PBX_REQUIRE(n_dims >= 1 && n_dims <= 128, "pbx_mvn_logpdf: n_dims must be in 1..128");
  PBX_REQUIRE(n_chains >= 0, "pbx_mvn_logpdf: negative point count");
  PBX_CUDA(cudaSetDevice(ctx->device));
  PBX_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  int rc = mvn_logpdf_launch(ctx, x, n_dims, n_chains, 1, mean, whiten, norm_c, log_pscale, out);
  if (rc) return rc;
  PBX_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  return PBX_OK;
}

template <int DQ, int NT>
static int gibbs_launch_nt(pbx_ctx* ctx, const GibbsArgs& a) {
  constexpr int DP = 4 * DQ;
  const size_t smem = ((size_t)DP * 4 * (DQ + 2) + 4 * DP) * sizeof(double);
  const int chains_per_cta = (NT / 32) * 8;
  const int grid = (a.C + chains_per_cta - 1) / chains_per_cta;
  const bool sweep_rec = (a.thin % a.d == 0) && (a.step0 % a.d == 0) && (a.T % a.d == 0);
  if (sweep_rec) {
    PBX_CUDA(cudaFuncSetAttribute(gibbs_mvn_kernel<DQ, true, NT>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gibbs_mvn_kernel<DQ, true, NT><<<grid, NT, smem, ctx->stream>>>(a);
  } else {
    PBX_CUDA(cudaFuncSetAttribute(gibbs_mvn_kernel<DQ, false, NT>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gibbs_mvn_kernel<DQ, false, NT><<<grid, NT, smem, ctx->stream>>>(a);
  }
  PBX_LAUNCH_CHECK(ctx);
  return PBX_OK;
}

template <int DQ>
static int gibbs_launch(pbx_ctx* ctx, const GibbsArgs& a) {
  // fewer than ~3 CTAs of 64 chains per SM: halve the CTA so that the grid covers the chip
  // (a strong-scaled shard of 8192 chains is 128 CTAs of 64 on 148 SMs)
  if ((int64_t)a.C < (int64_t)ctx->sm_count * 64 * 3) return gibbs_launch_nt<DQ, 128>(ctx, a);
  return gibbs_launch_nt<DQ, GB_THREADS>(ctx, a);
}

extern "C" int pbx_gibbs_mvn_run(pbx_ctx* ctx, const pbx_gibbs_mvn_params* p) {
  PBX_REQUIRE(ctx != nullptr && p != nullptr, "pbx_gibbs_mvn_run: null argument");
  PBX_REQUIRE(p->n_chains >= 1, "pbx_gibbs_mvn_run: n_chains must be >= 1");
  PBX_REQUIRE(p->n_dims >= 1 && p->n_dims <= 128, "pbx_gibbs_mvn_run: n_dims must be in 1..128");
  PBX_REQUIRE(p->n_steps >= 0 && p->thin >= 1, "pbx_gibbs_mvn_run: n_steps >= 0, thin >= 1");
  PBX_REQUIRE(p->step0 >= 0 && p->chain0 >= 0, "pbx_gibbs_mvn_run: step0/chain0 must be >= 0");
  PBX_REQUIRE(p->mean && p->coef && p->stdv && p->cdf_lo && p->cdf_hi && p->state,
              "pbx_gibbs_mvn_run: model/state pointers are mandatory");
  PBX_REQUIRE(!p->out_prob || (p->out_x && p->whiten),
              "pbx_gibbs_mvn_run: out_prob needs out_x and the whitening matrix");
  PBX_CUDA(cudaSetDevice(ctx->device));
  const int d = p->n_dims;
  int rc = pbx_ws_reserve(ctx, (size_t)d * 8);
  if (rc) return rc;
  double* c0 = (double*)ctx->ws;
  GibbsArgs a;
  a.C = p->n_chains; a.d = d; a.T = p->n_steps; a.thin = p->thin;
  a.step0 = p->step0; a.chain0 = p->chain0; a.seed = p->seed;
  a.coef = p->coef; a.c0 = c0; a.stdv = p->stdv; a.cdf_lo = p->cdf_lo; a.cdf_hi = p->cdf_hi;
  a.state = p->state; a.inj_runif = p->inj_runif; a.out_x = p->out_x;
  a.stat_sum = p->stat_sum; a.stat_sumsq = p->stat_sumsq;
  PBX_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  if (a.T > 0) {
    gibbs_c0_kernel<<<1, 128, 0, ctx->stream>>>(p->coef, p->mean, d, c0);
    PBX_LAUNCH_CHECK(ctx);
    if (d <= 4) rc = gibbs_launch<1>(ctx, a);
    else if (d <= 8) rc = gibbs_launch<2>(ctx, a);
    else if (d <= 16) rc = gibbs_launch<4>(ctx, a);
    else if (d <= 32) rc = gibbs_launch<8>(ctx, a);
    else if (d <= 64) rc = gibbs_launch<16>(ctx, a);
    else rc = gibbs_launch<32>(ctx, a);
    if (rc) return rc;
    if (p->out_prob && p->want_prob) {
      // the target is evaluated and recorded on every kept step (sd.py:286)
      rc = mvn_logpdf_launch(ctx, p->out_x, d, p->n_chains, p->n_steps / p->thin,
                             p->dens_mean ? p->dens_mean : p->mean, p->whiten, p->norm_c,
                             p->log_pscale, p->out_prob);
      if (rc) return rc;
    }
  }
  PBX_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  return PBX_OK;
}
