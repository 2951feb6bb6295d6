// pbx_grid.cu -- K3/K4: discrete grid exact inference of a normal (mu, sigma)
// posterior (examples/dgei/dgei_norm1d_improved.py:36-43).
//
// K3  log-joint[m][s] = (logprior_mu[m] + logprior_sigma[s])
//                       + sum_i norm.logpdf(x_i; mu_m, sigma_s)
//     The reference materialises the [N, M, S] tensor and np.sum(axis=0)s it
//     (probayes/rf.py:565-581, pd.py:368), then prod_rule adds the priors
//     (pd_utils.py:85-328, pscales.py:160-216).  Here every CTA owns 1 mu x 1024
//     sigma cells, keeps 8 running sums per thread in registers and streams the
//     observations through shared memory (1-D TMA bulk copies, 3-stage mbarrier
//     pipeline, 128-bit broadcast reads): 8 B * N of L2 traffic per 1024 cells, no
//     [N, M, S] temporary.  FP64-pipe bound: 2 instructions per (cell, observation).
// K4  PD.conditionalise (pd.py:285-295): p - max; exp; / max(tiny, sum); clamped log
//     PD.marginalise   (pd.py:162-164): exp (no shift); sum over an axis; clamped log
//     as deterministic two-stage reductions; HBM bound.
#include <math.h>
#include "pbx_common.cuh"

#define GR_TILE 1024
#define GR_STAGES 3
#define GR_THREADS 128
#define GR_KS 8                        // sigma cells per thread

__global__ void __launch_bounds__(GR_THREADS)
    grid_logjoint_kernel(const double* __restrict__ x, int64_t N, const double* __restrict__ mu,
                         const double* __restrict__ sigma, int S,
                         const double* __restrict__ lp_mu, const double* __restrict__ lp_sigma,
                         double* __restrict__ out, int use_tma) {
  __shared__ __align__(128) double sm[GR_STAGES][GR_TILE];
  __shared__ __align__(8) unsigned long long full_bar[GR_STAGES];
  const int m = blockIdx.y;
  const int s_base = blockIdx.x * (GR_THREADS * GR_KS) + threadIdx.x;
  const double mu_m = mu[m];
  double isg[GR_KS], acc[GR_KS];
#pragma unroll
  for (int k = 0; k < GR_KS; ++k) {
    const int s = s_base + k * GR_THREADS;
    isg[k] = (s < S) ? 1.0 / sigma[s] : 0.0;
    acc[k] = 0.0;
  }
  const int64_t n_full = N / GR_TILE;
  constexpr uint32_t kTileBytes = GR_TILE * sizeof(double);
  if (threadIdx.x == 0) {
    for (int s = 0; s < GR_STAGES; ++s) pbx_mbar_init(&full_bar[s], 1);
    pbx_fence_barrier_init();
  }
  __syncthreads();
  auto tile_math = [&](const double* sx, int cnt) {
#pragma unroll 2
    for (int i = 0; i < cnt; i += 2) {
      const double2 xv = *reinterpret_cast<const double2*>(sx + i);
      const double d0 = xv.x - mu_m, d1 = xv.y - mu_m;
#pragma unroll
      for (int k = 0; k < GR_KS; ++k) {
        const double z0 = d0 * isg[k], z1 = d1 * isg[k];
        acc[k] = fma(z0, z0, acc[k]);
        acc[k] = fma(z1, z1, acc[k]);
      }
    }
  };
  if (use_tma) {
    auto issue = [&](int64_t t) {
      const int s = (int)(t % GR_STAGES);
      pbx_mbar_expect_tx(&full_bar[s], kTileBytes);
      pbx_bulk_g2s(sm[s], x + t * GR_TILE, kTileBytes, &full_bar[s]);
    };
    if (threadIdx.x == 0)
      for (int64_t t = 0; t < n_full && t < GR_STAGES; ++t) issue(t);
    for (int64_t t = 0; t < n_full; ++t) {
      const int s = (int)(t % GR_STAGES);
      pbx_mbar_wait(&full_bar[s], (uint32_t)(t / GR_STAGES) & 1);
      tile_math(sm[s], GR_TILE);
      __syncthreads();
      if (threadIdx.x == 0 && t + GR_STAGES < n_full) issue(t + GR_STAGES);
    }
  } else {
    for (int64_t t = 0; t < n_full; ++t) {
      for (int i = threadIdx.x; i < GR_TILE; i += GR_THREADS) sm[0][i] = x[t * GR_TILE + i];
      __syncthreads();
      tile_math(sm[0], GR_TILE);
      __syncthreads();
    }
  }
  const int rem = (int)(N - n_full * GR_TILE);
  if (rem > 0) {
    __syncthreads();
    for (int i = threadIdx.x; i < GR_TILE; i += GR_THREADS)
      sm[0][i] = (i < rem) ? x[n_full * GR_TILE + i] : 0.0;
    __syncthreads();
    const int even = rem & ~1;
    tile_math(sm[0], even);
    if (rem & 1) {
      const double d0 = sm[0][even] - mu_m;
#pragma unroll
      for (int k = 0; k < GR_KS; ++k) {
        const double z0 = d0 * isg[k];
        acc[k] = fma(z0, z0, acc[k]);
      }
    }
  }
  const double n = (double)N, lpm = lp_mu[m];
#pragma unroll
  for (int k = 0; k < GR_KS; ++k) {
    const int s = s_base + k * GR_THREADS;
    if (s < S) {
      const double ll = -acc[k] * 0.5 - n * (PBX_LOG_SQRT_2PI + log(sigma[s]));
      out[(int64_t)m * S + s] = (lpm + lp_sigma[s]) + ll;      // prior first (prod_rule)
    }
  }
}

// ---------------------------------------------------------------------------
// deterministic two-stage reductions over n entries
// ---------------------------------------------------------------------------
#define RD_THREADS 256
enum { RD_MAX = 0, RD_SUMEXP = 1 };

template <int OP>
__device__ __forceinline__ double rd_combine(double a, double b) {
  return OP == RD_MAX ? fmax(a, b) : a + b;
}

template <int OP>
__global__ void __launch_bounds__(RD_THREADS)
    grid_reduce_stage1(const double* __restrict__ v, int64_t n, const double* __restrict__ gmax,
                       double* __restrict__ partial) {
  const double shift = OP == RD_SUMEXP ? gmax[0] : 0.0;
  double acc = OP == RD_MAX ? -INFINITY : 0.0, acc2 = acc;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if ((((uintptr_t)v) & 15) == 0) {                     // 128-bit loads, 4 in flight per thread
    const double2* v2 = reinterpret_cast<const double2*>(v);
    const int64_t n2 = n / 2;
#pragma unroll 4
    for (int64_t i = tid; i < n2; i += stride) {
      const double2 e = __ldcs(v2 + i);
      acc = rd_combine<OP>(acc, OP == RD_MAX ? e.x : pbx_exp_logp(e.x - shift));
      acc2 = rd_combine<OP>(acc2, OP == RD_MAX ? e.y : pbx_exp_logp(e.y - shift));
    }
    if ((n & 1) && tid == 0)
      acc = rd_combine<OP>(acc, OP == RD_MAX ? v[n - 1] : pbx_exp_logp(v[n - 1] - shift));
  } else {
    for (int64_t i = tid; i < n; i += stride) {
      const double e = v[i];
      acc = rd_combine<OP>(acc, OP == RD_MAX ? e : pbx_exp_logp(e - shift));
    }
  }
  acc = rd_combine<OP>(acc, acc2);
  __shared__ double sh[RD_THREADS];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = RD_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] = rd_combine<OP>(sh[threadIdx.x], sh[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

template <int OP>
__global__ void __launch_bounds__(RD_THREADS)
    grid_reduce_stage2(const double* __restrict__ partial, int np, double* __restrict__ out) {
  double acc = OP == RD_MAX ? -INFINITY : 0.0;
  for (int i = threadIdx.x; i < np; i += RD_THREADS) acc = rd_combine<OP>(acc, partial[i]);
  __shared__ double sh[RD_THREADS];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = RD_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] = rd_combine<OP>(sh[threadIdx.x], sh[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}

// ---------------------------------------------------------------------------
// posterior + marginals: CTA = PR_ROWS rows x 256 columns, thread = one column
// ---------------------------------------------------------------------------
#define PR_ROWS 32
#define PR_COLS 256

__global__ void __launch_bounds__(PR_COLS)
    grid_posterior_kernel(const double* lj, int M, int S,
                          const double* __restrict__ gmax, const double* __restrict__ gsum,
                          double* post, double* __restrict__ row_partial,
                          double* __restrict__ col_partial, int n_colblocks) {
  __shared__ double s_row[PR_ROWS][PR_COLS / 32];
  const int s = blockIdx.x * PR_COLS + threadIdx.x;
  const int m0 = blockIdx.y * PR_ROWS;
  const double mx = gmax[0];
  const double den = fmax(PBX_TINY, gsum[0]);
  const double lden = log(den);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double col = 0.0;
  for (int r = 0; r < PR_ROWS; ++r) {
    const int m = m0 + r;
    double q = 0.0;
    if (m < M && s < S) {
      // reference: q = exp_logp(lj - max) / max(tiny, sum); post = log_prob(q);
      // marginal term = exp_logp(post).  log(q) is evaluated as (lj - max) - log(den)
      // and exp(log q) as q itself: same values to ~1e-16 relative, a third of the
      // transcendental work; the clamp decision (q < tiny -> -1.797e308) is the
      // reference's, taken on q.
      const double sh = lj[(int64_t)m * S + s] - mx;
      const double e = pbx_exp_logp(sh);
      q = e / den;
      const bool keep = q >= PBX_TINY;
      if (post) post[(int64_t)m * S + s] = keep ? sh - lden : -PBX_HUGE;
      q = keep ? q : 0.0;                                 // exp_logp(-1.797e308) = 0
    }
    col += q;
    double w = q;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if (lane == 0) s_row[r][warp] = w;
  }
  if (s < S) col_partial[(int64_t)blockIdx.y * S + s] = col;
  __syncthreads();
  if (threadIdx.x < PR_ROWS) {
    const int m = m0 + threadIdx.x;
    if (m < M) {
      double w = 0.0;
#pragma unroll
      for (int k = 0; k < PR_COLS / 32; ++k) w += s_row[threadIdx.x][k];
      row_partial[(int64_t)m * n_colblocks + blockIdx.x] = w;
    }
  }
}

__global__ void __launch_bounds__(256)
    grid_marginal_finish(const double* __restrict__ row_partial, int M, int n_colblocks,
                         const double* __restrict__ col_partial, int n_rowblocks, int S,
                         double* __restrict__ marg_mu, double* __restrict__ marg_sigma) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M && marg_mu) {
    double w = 0.0;
    for (int k = 0; k < n_colblocks; ++k) w += row_partial[(int64_t)i * n_colblocks + k];
    marg_mu[i] = w;
  }
  if (i < S && marg_sigma) {
    double w = 0.0;
    for (int k = 0; k < n_rowblocks; ++k) w += col_partial[(int64_t)k * S + i];
    marg_sigma[i] = w;
  }
}

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" int pbx_grid_norm_logjoint(pbx_ctx* ctx, const double* x_obs, int64_t n_obs,
                                      const double* mu, int32_t n_mu, const double* sigma,
                                      int32_t n_sigma, const double* logprior_mu,
                                      const double* logprior_sigma, double* out) {
  PBX_REQUIRE(ctx && x_obs && mu && sigma && logprior_mu && logprior_sigma && out,
              "pbx_grid_norm_logjoint: null argument");
  PBX_REQUIRE(n_obs >= 1 && n_mu >= 1 && n_sigma >= 1,
              "pbx_grid_norm_logjoint: sizes must be positive");
  PBX_REQUIRE(n_mu <= 65535, "pbx_grid_norm_logjoint: at most 65535 mu rows per call (slab it)");
  PBX_CUDA(cudaSetDevice(ctx->device));
  dim3 grid((n_sigma + GR_THREADS * GR_KS - 1) / (GR_THREADS * GR_KS), n_mu);
  const int use_tma = (((uintptr_t)x_obs) % 16 == 0) ? 1 : 0;
  PBX_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  grid_logjoint_kernel<<<grid, GR_THREADS, 0, ctx->stream>>>(
      x_obs, n_obs, mu, sigma, n_sigma, logprior_mu, logprior_sigma, out, use_tma);
  PBX_LAUNCH_CHECK(ctx);
  PBX_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  return PBX_OK;
}

// ---------------------------------------------------------------------------
// opt-in: the same log-joint from centred sufficient statistics of the observations
// (see pbx_mh_normreg.cu, variant 3): sum (x - mu)^2 = RSS + N a^2 - 2 a Se with
// a = mu - mean(x), every term non-negative or a rounding residual.  O(M S) instead of
// O(N M S): 4096^2 cells in ~50 us instead of 198 ms, same values to <= 1e-15.
// ---------------------------------------------------------------------------
#define GSS_ROWS 32
__global__ void __launch_bounds__(256) grid_logjoint_ss_kernel(
    const NrStats* __restrict__ stp, const double* __restrict__ mu,
    const double* __restrict__ sigma, int M, int S, const double* __restrict__ lp_mu,
    const double* __restrict__ lp_sigma, double* __restrict__ out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const NrStats st = *stp;
  // per-column terms once, then GSS_ROWS cells of the column (log and the reciprocal are
  // the expensive part of a cell)
  const double sg = sigma[s], isg = 1.0 / sg;
  const double colc = st.N * (PBX_LOG_SQRT_2PI + log(sg)), lps = lp_sigma[s];
  const int m1 = min(M, (int)(blockIdx.y + 1) * GSS_ROWS);
  for (int m = blockIdx.y * GSS_ROWS; m < m1; ++m) {
    const double a = mu[m] - st.cy;
    double ssq = fma(st.N * a, a, st.RSS);
    ssq = fma(-2.0 * a, st.Se, ssq);
    const double acc = (ssq * isg) * isg;
    const double ll = -acc * 0.5 - colc;
    __stcs(out + (int64_t)m * S + s, (lp_mu[m] + lps) + ll);
  }
}

extern "C" int pbx_grid_norm_logjoint_ss(pbx_ctx* ctx, const double* x_obs, int64_t n_obs,
                                         const double* mu, int32_t n_mu, const double* sigma,
                                         int32_t n_sigma, const double* logprior_mu,
                                         const double* logprior_sigma, double* out) {
  PBX_REQUIRE(ctx && x_obs && mu && sigma && logprior_mu && logprior_sigma && out,
              "pbx_grid_norm_logjoint_ss: null argument");
  PBX_REQUIRE(n_obs >= 1 && n_mu >= 1 && n_sigma >= 1,
              "pbx_grid_norm_logjoint_ss: sizes must be positive");
  PBX_REQUIRE(n_mu <= 65535, "pbx_grid_norm_logjoint_ss: at most 65535 mu rows per call (slab it)");
  PBX_CUDA(cudaSetDevice(ctx->device));
  PBX_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  NrStats* st = nullptr;
  int rc = pbx_ss_compute(ctx, nullptr, x_obs, n_obs, &st);
  if (rc) return rc;
  dim3 grid((n_sigma + 255) / 256, (n_mu + GSS_ROWS - 1) / GSS_ROWS);
  grid_logjoint_ss_kernel<<<grid, 256, 0, ctx->stream>>>(st, mu, sigma, n_mu, n_sigma,
                                                         logprior_mu, logprior_sigma, out);
  PBX_LAUNCH_CHECK(ctx);
  PBX_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  return PBX_OK;
}

template <int OP>
static int grid_reduce(pbx_ctx* ctx, const double* v, int64_t n, const double* gmax, double* out) {
  PBX_CUDA(cudaSetDevice(ctx->device));
  int64_t want = (n + RD_THREADS * 16 - 1) / (RD_THREADS * 16);
  int np = (int)(want < 1 ? 1 : (want > ctx->sm_count * 8 ? ctx->sm_count * 8 : want));
  int rc = pbx_ws_reserve(ctx, (size_t)np * 8);
  if (rc) return rc;
  double* partial = (double*)ctx->ws;
  PBX_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  grid_reduce_stage1<OP><<<np, RD_THREADS, 0, ctx->stream>>>(v, n, gmax, partial);
  PBX_LAUNCH_CHECK(ctx);
  grid_reduce_stage2<OP><<<1, RD_THREADS, 0, ctx->stream>>>(partial, np, out);
  PBX_LAUNCH_CHECK(ctx);
  PBX_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  return PBX_OK;
}

extern "C" int pbx_grid_max(pbx_ctx* ctx, const double* logjoint, int64_t n, double* out) {
  PBX_REQUIRE(ctx && logjoint && out && n >= 1, "pbx_grid_max: bad argument");
  return grid_reduce<RD_MAX>(ctx, logjoint, n, nullptr, out);
}

extern "C" int pbx_grid_sumexp(pbx_ctx* ctx, const double* logjoint, int64_t n,
                               const double* gmax, double* out) {
  PBX_REQUIRE(ctx && logjoint && gmax && out && n >= 1, "pbx_grid_sumexp: bad argument");
  return grid_reduce<RD_SUMEXP>(ctx, logjoint, n, gmax, out);
}

extern "C" int pbx_grid_posterior(pbx_ctx* ctx, const double* logjoint, int32_t n_mu,
                                  int32_t n_sigma, const double* gmax, const double* gsum,
                                  double* post, double* marg_mu_lin, double* marg_sigma_lin) {
  PBX_REQUIRE(ctx && logjoint && gmax && gsum, "pbx_grid_posterior: null argument");
  PBX_REQUIRE(n_mu >= 1 && n_sigma >= 1, "pbx_grid_posterior: sizes must be positive");
  PBX_CUDA(cudaSetDevice(ctx->device));
  const int ncb = (n_sigma + PR_COLS - 1) / PR_COLS, nrb = (n_mu + PR_ROWS - 1) / PR_ROWS;
  PBX_REQUIRE(nrb <= 65535, "pbx_grid_posterior: too many rows per call (slab it)");
  const size_t rp = ((size_t)n_mu * ncb * 8 + 255) / 256 * 256;
  const size_t cp = (size_t)nrb * n_sigma * 8;
  int rc = pbx_ws_reserve(ctx, rp + cp);
  if (rc) return rc;
  double* row_partial = (double*)ctx->ws;
  double* col_partial = (double*)((char*)ctx->ws + rp);
  PBX_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  grid_posterior_kernel<<<dim3(ncb, nrb), PR_COLS, 0, ctx->stream>>>(
      logjoint, n_mu, n_sigma, gmax, gsum, post, row_partial, col_partial, ncb);
  PBX_LAUNCH_CHECK(ctx);
  if (marg_mu_lin || marg_sigma_lin) {
    const int n = n_mu > n_sigma ? n_mu : n_sigma;
    grid_marginal_finish<<<(n + 255) / 256, 256, 0, ctx->stream>>>(
        row_partial, n_mu, ncb, col_partial, nrb, n_sigma, marg_mu_lin, marg_sigma_lin);
    PBX_LAUNCH_CHECK(ctx);
  }
  PBX_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  return PBX_OK;
}

// ---------------------------------------------------------------------------
// K4 in two passes over the grid (was max, sum-exp, posterior: three reads + a write).
//
// Pass A  pbx_grid_max_sumexp: ONE read.  Every thread keeps an online (max, sum-exp)
//         pair: per chunk of 8 entries held in registers, the running maximum is raised
//         first (one rescaling exp per chunk, and only when the maximum moves), then the
//         8 terms exp(v - max) are added.  Pairs are merged (m, s) + (m', s') =
//         (M = max, s exp(m - M) + s' exp(m' - M)) by shuffles and shared memory in a
//         fixed order; the last CTA to finish (ticket) merges the per-CTA pairs in index
//         order.  The result differs from sum exp(v - global max) only by the rounding of
//         the rescaling factors (~1e-16 relative).  Linear-pscale input: plain sum.
// Pass B  pbx_grid_posterior2: one read + one write (or in place), both marginals from
//         the same pass, their clamped logs fused into the finishing kernel.
// ---------------------------------------------------------------------------
// Table-driven exp for the two passes (libm's exp costs ~70 SASS instructions and made pass
// A instruction-bound at 2.2 TB/s): exp(x) = 2^(n/64) 2^k e^r, |r| <= ln2/128, one 8-byte
// table entry + a degree-5 polynomial, <= 2 ulp on [-700, 700] (libm outside: a rare branch).
// The 64-entry table is stored 16 times interleaved (entry j of copy g at 16 j + g) and lane
// l reads copy l % 16, so the per-lane lookups never collide in a shared-memory bank; each
// CTA stages the 8 KB into shared memory.
static __device__ double g_exptab[64 * 16];
// (constants live in the constant bank: FP64 instructions take no 64-bit immediates, and a
// literal costs two MOVs at every use)
__constant__ double kGE[12] = {92.332482616893656877,     // 0: 64 / ln2
                               6755399441055744.0,        // 1: 1.5 * 2^52
                               -0.01083042469326756,      // 2: -ln2/64 hi (32 bits)
                               -2.9815858269852933e-12,   // 3: -ln2/64 lo
                               1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5,   // 4..7
                               -600.0, 100.0, -700.0,     // 8..10
                               3.7200759760208361e-44};   // 11: exp(-100)
__device__ __forceinline__ double grid_fast_exp(double x, const double* tab) {
  const double fn = fma(x, kGE[0], kGE[1]);
  const int n = __double2loint(fn);
  const double k = fn - kGE[1];
  double r = fma(k, kGE[2], x);                        // k * hi is exact
  r = fma(k, kGE[3], r);
  double p = fma(r, kGE[4], kGE[5]);
  p = fma(r, p, kGE[6]);
  p = fma(r, p, kGE[7]);
  p = fma(r * r, p, r);                                                    // expm1(r)
  const double tj = tab[((n & 63) << 4) | (threadIdx.x & 15)];
  const double y = fma(tj, p, tj);
  return __hiloint2double(__double2hiint(y) + ((n >> 6) << 20), __double2loint(y));
}
// clamped exp of pscales.py:56-65 over the whole double range, BRANCH-FREE (a branch per
// element kept the compiler from interleaving the elements of a chunk): below -600 the
// argument is shifted by +100 and the result multiplied by exp(-100), so that results down to
// the subnormals come out of one correctly rounded multiply; above log(huge) the reference
// returns huge.
__device__ __forceinline__ double grid_exp_logp(double l, const double* tab) {
  const bool lowx = l < kGE[8];
  const double xs = lowx ? fmax(l + kGE[9], kGE[10]) : fmin(l, PBX_LOG_HUGE);
  double y = grid_fast_exp(xs, tab);
  y = lowx ? y * kGE[11] : y;                              // exp(-100)
  return (l <= PBX_LOG_HUGE) ? y : PBX_HUGE;               // NaN -> huge, as the reference
}
__device__ __forceinline__ void grid_stage_exptab(double* s_tab, int nthreads) {
  for (int i = threadIdx.x; i < 64 * 16; i += nthreads) s_tab[i] = g_exptab[i];
  __syncthreads();
}
static int grid_init_exptab(pbx_ctx* ctx) {
  if (ctx->exptab_ready) return PBX_OK;
  double* h = new double[64 * 16];
  for (int j = 0; j < 64; ++j)
    for (int g = 0; g < 16; ++g) h[16 * j + g] = (double)exp2l(j / 64.0L);
  cudaError_t e = cudaMemcpyToSymbolAsync(g_exptab, h, sizeof(double) * 64 * 16, 0,
                                          cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  delete[] h;
  PBX_CUDA(e);
  ctx->exptab_ready = true;
  return PBX_OK;
}

struct MsPair { double m, s; };
__device__ __forceinline__ MsPair ms_merge(MsPair a, MsPair b) {
  const double M = fmax(a.m, b.m);
  MsPair r;
  r.m = M;
  // exp(-inf - -inf) never occurs: a pair with m = -inf has s = 0 and is skipped
  r.s = (a.s == 0.0 ? 0.0 : a.s * exp(a.m - M)) + (b.s == 0.0 ? 0.0 : b.s * exp(b.m - M));
  return r;
}

#define MS_THREADS 256
#define MS_CHUNK 8
template <bool kLinear>
__global__ void __launch_bounds__(MS_THREADS)
    grid_max_sumexp_kernel(const double* __restrict__ v, int64_t n, MsPair* __restrict__ partial,
                           unsigned int* __restrict__ ticket, double* __restrict__ out2) {
  __shared__ double s_tab[64 * 16];
  if (!kLinear) grid_stage_exptab(s_tab, MS_THREADS);
  double m = -INFINITY, s = 0.0;
  // a warp owns 256 consecutive entries per turn: load i of lane l is the double2 at
  // 32 i + l of the block, so every 128-bit load instruction is one fully used 512-byte
  // segment (a per-thread run of 64 bytes leaves half of each sector to the next load)
  constexpr int WCH = 32 * MS_CHUNK;
  const int64_t nwch = n / WCH;
  const int lane_ = threadIdx.x & 31;
  const int64_t wstride = (int64_t)gridDim.x * (MS_THREADS / 32);
  const bool vec = (((uintptr_t)v) & 15) == 0;
  auto fold = [&](const double (&e)[MS_CHUNK], unsigned mask) {      // bit i: slot i holds an entry
    if (kLinear) {
#pragma unroll
      for (int i = 0; i < MS_CHUNK; ++i) s += ((mask >> i) & 1u) ? e[i] : 0.0;
    } else {
      double cm = -INFINITY;
#pragma unroll
      for (int i = 0; i < MS_CHUNK; ++i) cm = fmax(cm, ((mask >> i) & 1u) ? e[i] : -INFINITY);
      if (cm > m) {
        s = (s == 0.0) ? 0.0 : s * exp(m - cm);
        m = cm;
      }
      double t0 = 0.0, t1 = 0.0;
#pragma unroll
      for (int i = 0; i < MS_CHUNK; i += 2) {
        // e - m <= 0 here; below -700 a term is < 1e-304 of a sum that is >= 1 (the maximum
        // contributes exp(0)): clamping the argument changes no bit of the result
        t0 += ((mask >> i) & 1u) ? grid_fast_exp(fmax(e[i] - m, kGE[10]), s_tab) : 0.0;
        t1 += ((mask >> (i + 1)) & 1u) ? grid_fast_exp(fmax(e[i + 1] - m, kGE[10]), s_tab) : 0.0;
      }
      s += t0 + t1;
    }
  };
  for (int64_t wc = (int64_t)blockIdx.x * (MS_THREADS / 32) + (threadIdx.x >> 5); wc < nwch;
       wc += wstride) {
    double e[MS_CHUNK];
    const double* base = v + wc * WCH;
    if (vec) {
      const double2* v2 = reinterpret_cast<const double2*>(base);
#pragma unroll
      for (int i = 0; i < MS_CHUNK / 2; ++i) {
        const double2 t = __ldcs(v2 + 32 * i + lane_);
        e[2 * i] = t.x;
        e[2 * i + 1] = t.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < MS_CHUNK; ++i) e[i] = base[32 * i + lane_];
    }
    fold(e, (1u << MS_CHUNK) - 1u);
  }
  if (blockIdx.x == 0 && threadIdx.x < 32) {               // the n % 256 tail: one warp
    double e[MS_CHUNK];
    unsigned mask = 0;
#pragma unroll
    for (int i = 0; i < MS_CHUNK; ++i) {
      const int64_t idx = nwch * WCH + 32 * i + lane_;
      e[i] = (idx < n) ? v[idx] : 0.0;
      mask |= (idx < n) ? (1u << i) : 0u;
    }
    if (mask) fold(e, mask);
  }
  MsPair p;
  p.m = kLinear ? 0.0 : m;
  p.s = s;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    MsPair q;
    q.m = __shfl_xor_sync(0xffffffffu, p.m, o);
    q.s = __shfl_xor_sync(0xffffffffu, p.s, o);
    if (kLinear) p.s += q.s;
    else p = ms_merge(p, q);              // symmetric: both lanes of a pair get the same bits
  }
  __shared__ MsPair sh[MS_THREADS / 32];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sh[warp] = p;
  __syncthreads();
  if (threadIdx.x == 0) {
    MsPair a = sh[0];
    for (int w = 1; w < MS_THREADS / 32; ++w) {
      if (kLinear) a.s += sh[w].s;
      else a = ms_merge(a, sh[w]);
    }
    partial[blockIdx.x] = a;
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {                                              // fixed tree: deterministic
    __threadfence();
    const volatile MsPair* pv = partial;
    MsPair a;
    a.m = kLinear ? 0.0 : -INFINITY;
    a.s = 0.0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += MS_THREADS) {
      MsPair b;
      b.m = pv[i].m;
      b.s = pv[i].s;
      if (kLinear) a.s += b.s;
      else a = ms_merge(a, b);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      MsPair q;
      q.m = __shfl_xor_sync(0xffffffffu, a.m, o);
      q.s = __shfl_xor_sync(0xffffffffu, a.s, o);
      if (kLinear) a.s += q.s;
      else a = ms_merge(a, q);
    }
    __syncthreads();                                       // sh[] is free again
    if (lane == 0) sh[warp] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
      MsPair t = sh[0];
      for (int w = 1; w < MS_THREADS / 32; ++w) {
        if (kLinear) t.s += sh[w].s;
        else t = ms_merge(t, sh[w]);
      }
      out2[0] = t.m;
      out2[1] = t.s;
      *ticket = 0;                                         // ready for the next call
    }
  }
}

__global__ void grid_rescale_sumexp_kernel(const double* lmax, const double* gmax, double* sum) {
  sum[0] = (sum[0] == 0.0) ? 0.0 : sum[0] * exp(lmax[0] - gmax[0]);
}

#define P2_ROWS 32
#define P2_CPT 4                        // columns per thread (two 128-bit accesses per row)
#define P2_THREADS 256
#define P2_COLS (P2_THREADS * P2_CPT)
#define P2_RU 4                         // rows in flight per thread

// kLinear: the input (and the output) are linear-pscale probabilities: post = p / max(tiny,
// sum) (pd.py:285-295 without the log/exp round trip), marginals = plain sums.
template <bool kLinear>
__global__ void __launch_bounds__(P2_THREADS)
    grid_posterior2_kernel(const double* lj, int M, int S, const double* __restrict__ gmax,
                           const double* __restrict__ gsum, double* post,
                           double* __restrict__ row_partial, double* __restrict__ col_partial,
                           int n_colblocks, double qmin) {
  __shared__ double s_row[P2_ROWS][P2_THREADS / 32];
  __shared__ double s_tab[64 * 16];
  if (!kLinear) grid_stage_exptab(s_tab, P2_THREADS);
  // Column ownership: a warp covers 128 consecutive columns; thread (warp w, lane l) owns
  // the two column PAIRS at cb + 2 l and cb + 64 + 2 l, cb = block base + 128 w, so each of
  // its two 128-bit accesses per row is, warp-wide, one contiguous 512-byte segment.
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cb = blockIdx.x * P2_COLS + warp * 128 + 2 * lane;
  const int m0 = blockIdx.y * P2_ROWS;
  const double mx = kLinear ? 0.0 : gmax[0];
  const double den = fmax(PBX_TINY, gsum[0]);
  const double lden = kLinear ? 0.0 : log(den);
  const double rden = 1.0 / den;
  // 128-bit path: even row length, 16-byte aligned bases
  const bool vec = (S % 2 == 0) && ((((uintptr_t)lj) & 15) == 0) &&
                   (post == nullptr || (((uintptr_t)post) & 15) == 0);
  int colidx[P2_CPT];
  colidx[0] = cb; colidx[1] = cb + 1; colidx[2] = cb + 64; colidx[3] = cb + 65;
  double col[P2_CPT];
#pragma unroll
  for (int k = 0; k < P2_CPT; ++k) col[k] = 0.0;
  for (int r0 = 0; r0 < P2_ROWS; r0 += P2_RU) {
    double e[P2_RU][P2_CPT];
    // all loads of the P2_RU rows first (post may alias lj: the compiler cannot hoist them)
#pragma unroll
    for (int u = 0; u < P2_RU; ++u) {
      const int m = m0 + r0 + u;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c0 = colidx[2 * h];
        if (m < M && vec && c0 + 1 < S) {
          const double2 a = __ldcs(reinterpret_cast<const double2*>(lj + (int64_t)m * S + c0));
          e[u][2 * h] = a.x;
          e[u][2 * h + 1] = a.y;
        } else {
#pragma unroll
          for (int k = 0; k < 2; ++k)
            e[u][2 * h + k] = (m < M && c0 + k < S) ? lj[(int64_t)m * S + c0 + k]
                                                    : (kLinear ? 0.0 : -PBX_HUGE);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < P2_RU; ++u) {
      const int m = m0 + r0 + u;
      double q[P2_CPT], o[P2_CPT];
#pragma unroll
      for (int k = 0; k < P2_CPT; ++k) {
        if (kLinear) {
          // reference: div_prob(p, sum) = p / max(tiny, sum)
          q[k] = e[u][k] / den;
          o[k] = q[k];
        } else {
          // reference: q = exp_logp(lj - max) / max(tiny, sum); post = log_prob(q); the
          // marginal term is exp_logp(post).  log(q) is evaluated as (lj - max) - log(den)
          // and exp(log q) as q itself: the same values to ~1e-16 relative for a third of
          // the transcendental work; the clamp decision (q < tiny -> -1.797e308) is the
          // reference's, taken on q.
          const double sh = e[u][k] - mx;
          const double qq = grid_exp_logp(sh, s_tab) * rden;  // q only feeds the marginal sums
          // (qmin = tiny; 0 for PD.marginalise, which sums exp_logp(p) unclamped)
          const bool keep = qq >= qmin;
          o[k] = keep ? sh - lden : -PBX_HUGE;
          q[k] = keep ? qq : 0.0;                           // exp_logp(-1.797e308) = 0
        }
        if (!(m < M && colidx[k] < S)) q[k] = 0.0;
        col[k] += q[k];
      }
      if (post && m < M) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c0 = colidx[2 * h];
          if (vec && c0 + 1 < S) {
            __stcs(reinterpret_cast<double2*>(post + (int64_t)m * S + c0),
                   make_double2(o[2 * h], o[2 * h + 1]));
          } else {
#pragma unroll
            for (int k = 0; k < 2; ++k)
              if (c0 + k < S) post[(int64_t)m * S + c0 + k] = o[2 * h + k];
          }
        }
      }
      double w = (q[0] + q[1]) + (q[2] + q[3]);
#pragma unroll
      for (int ofs = 16; ofs > 0; ofs >>= 1) w += __shfl_xor_sync(0xffffffffu, w, ofs);
      if (lane == 0) s_row[r0 + u][warp] = w;
    }
  }
#pragma unroll
  for (int k = 0; k < P2_CPT; ++k)
    if (colidx[k] < S) col_partial[(int64_t)blockIdx.y * S + colidx[k]] = col[k];
  __syncthreads();
  if (threadIdx.x < P2_ROWS) {
    const int m = m0 + threadIdx.x;
    if (m < M) {
      double w = 0.0;
#pragma unroll
      for (int k = 0; k < P2_THREADS / 32; ++k) w += s_row[threadIdx.x][k];
      row_partial[(int64_t)m * n_colblocks + blockIdx.x] = w;
    }
  }
}

// log_flags: bit 0 -> marg_mu through the clamped log, bit 1 -> marg_sigma
__global__ void __launch_bounds__(256)
    grid_marginal_finish2(const double* __restrict__ row_partial, int M, int n_colblocks,
                          const double* __restrict__ col_partial, int n_rowblocks, int S,
                          double* __restrict__ marg_mu, double* __restrict__ marg_sigma,
                          int log_flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M && marg_mu) {
    double w = 0.0;
    for (int k = 0; k < n_colblocks; ++k) w += row_partial[(int64_t)i * n_colblocks + k];
    marg_mu[i] = (log_flags & 1) ? pbx_log_prob(w) : w;
  }
  if (i < S && marg_sigma) {
    double w = 0.0;
    for (int k = 0; k < n_rowblocks; ++k) w += col_partial[(int64_t)k * S + i];
    marg_sigma[i] = (log_flags & 2) ? pbx_log_prob(w) : w;
  }
}

extern "C" int pbx_grid_max_sumexp(pbx_ctx* ctx, const double* v, int64_t n, int32_t linear,
                                   double* out2) {
  PBX_REQUIRE(ctx && v && out2 && n >= 1, "pbx_grid_max_sumexp: bad argument");
  PBX_CUDA(cudaSetDevice(ctx->device));
  {
    int rc0 = grid_init_exptab(ctx);
    if (rc0) return rc0;
  }
  int64_t want = (n + (int64_t)MS_THREADS * MS_CHUNK * 4 - 1) / ((int64_t)MS_THREADS * MS_CHUNK * 4);
  const int np = (int)(want < 1 ? 1 : (want > ctx->sm_count * 8 ? ctx->sm_count * 8 : want));
  // workspace: [ticket (256 B, zero between calls)] [np pairs]
  int rc = pbx_ws_reserve(ctx, 256 + (size_t)np * sizeof(MsPair));
  if (rc) return rc;
  unsigned int* ticket = (unsigned int*)ctx->ws;
  MsPair* partial = (MsPair*)((char*)ctx->ws + 256);
  PBX_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  // the ticket word lives at the start of the shared workspace, which other entry points
  // overwrite: zero it every call (a 4-byte memset node, no kernel)
  PBX_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned int), ctx->stream));
  if (linear)
    grid_max_sumexp_kernel<true><<<np, MS_THREADS, 0, ctx->stream>>>(v, n, partial, ticket, out2);
  else
    grid_max_sumexp_kernel<false><<<np, MS_THREADS, 0, ctx->stream>>>(v, n, partial, ticket, out2);
  PBX_LAUNCH_CHECK(ctx);
  PBX_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  return PBX_OK;
}

extern "C" int pbx_grid_rescale_sumexp(pbx_ctx* ctx, const double* local_max,
                                       const double* global_max, double* sum_inout) {
  PBX_REQUIRE(ctx && local_max && global_max && sum_inout, "pbx_grid_rescale_sumexp: null argument");
  PBX_CUDA(cudaSetDevice(ctx->device));
  grid_rescale_sumexp_kernel<<<1, 1, 0, ctx->stream>>>(local_max, global_max, sum_inout);
  PBX_LAUNCH_CHECK(ctx);
  return PBX_OK;
}

extern "C" int pbx_grid_posterior2(pbx_ctx* ctx, const double* prob, int32_t n_mu, int32_t n_sigma,
                                   const double* gmax, const double* gsum, int32_t linear,
                                   double* post, double* marg_mu, double* marg_sigma,
                                   int32_t marg_log_flags) {
  PBX_REQUIRE(ctx && prob && gsum && (linear || gmax), "pbx_grid_posterior2: null argument");
  PBX_REQUIRE(n_mu >= 1 && n_sigma >= 1, "pbx_grid_posterior2: sizes must be positive");
  PBX_CUDA(cudaSetDevice(ctx->device));
  {
    int rc0 = grid_init_exptab(ctx);
    if (rc0) return rc0;
  }
  const int ncb = (n_sigma + P2_COLS - 1) / P2_COLS, nrb = (n_mu + P2_ROWS - 1) / P2_ROWS;
  PBX_REQUIRE(nrb <= 65535, "pbx_grid_posterior2: too many rows per call (slab it)");
  const size_t rp = ((size_t)n_mu * ncb * 8 + 255) / 256 * 256;
  const size_t cp = (size_t)nrb * n_sigma * 8;
  int rc = pbx_ws_reserve(ctx, rp + cp);
  if (rc) return rc;
  double* row_partial = (double*)ctx->ws;
  double* col_partial = (double*)((char*)ctx->ws + rp);
  PBX_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  const double qmin = (marg_log_flags & 4) ? 0.0 : PBX_TINY;
  if (linear)
    grid_posterior2_kernel<true><<<dim3(ncb, nrb), P2_THREADS, 0, ctx->stream>>>(
        prob, n_mu, n_sigma, gmax, gsum, post, row_partial, col_partial, ncb, qmin);
  else
    grid_posterior2_kernel<false><<<dim3(ncb, nrb), P2_THREADS, 0, ctx->stream>>>(
        prob, n_mu, n_sigma, gmax, gsum, post, row_partial, col_partial, ncb, qmin);
  PBX_LAUNCH_CHECK(ctx);
  if (marg_mu || marg_sigma) {
    const int n = n_mu > n_sigma ? n_mu : n_sigma;
    grid_marginal_finish2<<<(n + 255) / 256, 256, 0, ctx->stream>>>(
        row_partial, n_mu, ncb, col_partial, nrb, n_sigma, marg_mu, marg_sigma, marg_log_flags);
    PBX_LAUNCH_CHECK(ctx);
  }
  PBX_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  return PBX_OK;
}
