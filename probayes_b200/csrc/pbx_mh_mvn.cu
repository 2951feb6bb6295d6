// pbx_mh_mvn.cu -- K1: batched-chain Metropolis-Hastings on a multivariate-normal
// target.  Restates (per chain) the reference step loop probayes/sp.py:221-258 with
// the accept rule of sp_utils.py:19-37 / pscales.py:219-236 and the scipy mvn
// density the reference calls at prob.py:347-360.
//
// Layout: state[D][C], outputs [..][D][C] / [..][C] -- chain-minor so a warp of 32
// chains reads/writes 256 contiguous bytes per dimension.
//
// v1 kernel (this file): one thread = one chain, state + running sums in registers,
// Philox4x32-10 + Box-Muller in-thread (or injected streams for bit-parity runs),
// accept test and thinned write-back fused.
#include <math.h>
#include "pbx_common.cuh"

struct MhMvnConst {
  double mean[PBX_MAX_DIMS];
  double W[PBX_MAX_DIMS * PBX_MAX_DIMS];
  double L[PBX_MAX_DIMS * PBX_MAX_DIMS];
  double scale[PBX_MAX_DIMS];
  double norm_c;
};

struct MhMvnArgs {
  int C, T, thin;
  int64_t step0, chain0;
  uint64_t seed;
  int log_pscale, accept_mode, prop_kind, has_L;
  double* state;
  double* state_lp;
  const double* inj_delta;
  const double* inj_thresh;
  double* out_x;
  double* out_prob;
  uint8_t* out_accept;
  double* out_score;
  int64_t* accept_count;
  double* stat_sum;
  double* stat_sumsq;
};

// scipy multivariate_normal_gen._logpdf: -0.5*(rank*log(2pi) + log_pdet + maha),
// maha = sum(square(dev @ U)).
template <int D>
__device__ __forceinline__ double mvn_logpdf(const double (&x)[D], const MhMvnConst& m) {
  double dev[D];
#pragma unroll
  for (int j = 0; j < D; ++j) dev[j] = x[j] - m.mean[j];
  double maha = 0.0;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    double y = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) y = fma(dev[j], m.W[j * D + k], y);
    maha = fma(y, y, maha);
  }
  return -0.5 * (m.norm_c + maha);
}

template <int D, bool kInjected>
__global__ void __launch_bounds__(128) mh_mvn_kernel(const MhMvnArgs a,
                                                     const __grid_constant__ MhMvnConst m) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.C) return;
  const int64_t C = a.C;
  const uint32_t gchain = (uint32_t)(a.chain0 + c);

  double x[D], ssum[D], ssq[D];
#pragma unroll
  for (int j = 0; j < D; ++j) {
    x[j] = a.state[j * C + c];
    ssum[j] = 0.0;
    ssq[j] = 0.0;
  }
  double lp = (a.step0 > 0) ? a.state_lp[c] : 0.0;
  // linear density of the retained state (what the reference's opqr.o.prob holds)
  double lin = (a.step0 > 0) ? (a.log_pscale ? pbx_exp_logp(lp) : exp(lp)) : 0.0;
  int64_t nacc = 0;
  const bool ref_mode = (a.accept_mode == PBX_ACCEPT_REFERENCE);

  for (int k = 0; k < a.T; ++k) {
    const int64_t gstep = a.step0 + k;
    double dl[D];
    double t;
    // ---- draws: proposal first, then the threshold (sp.py:231-249) ----------
    if (kInjected) {
#pragma unroll
      for (int j = 0; j < D; ++j) dl[j] = a.inj_delta[((int64_t)k * D + j) * C + c];
      t = a.inj_thresh[(int64_t)k * C + c];
    } else {
#pragma unroll
      for (int s = 0; s < (D + 1) / 2; ++s) {
        pbx_u4 w = pbx_block(a.seed, (uint64_t)gstep, gchain, (uint32_t)s);
        double d0, d1;
        if (a.prop_kind == PBX_PROP_NORMAL) {
          pbx_normal_pair(w, d0, d1);
          d0 *= m.scale[2 * s];
          if (2 * s + 1 < D) d1 *= m.scale[2 * s + 1];
        } else {
          double r0 = pbx_u01(w.x, w.y), r1 = pbx_u01(w.z, w.w);
          d0 = -m.scale[2 * s] + (2.0 * m.scale[2 * s]) * r0;
          d1 = (2 * s + 1 < D) ? -m.scale[2 * s + 1] + (2.0 * m.scale[2 * s + 1]) * r1 : 0.0;
        }
        dl[2 * s] = d0;
        if (2 * s + 1 < D) dl[2 * s + 1] = d1;
      }
      pbx_u4 w = pbx_block(a.seed, (uint64_t)gstep, gchain, PBX_SLOT_THRESH);
      t = pbx_u01(w.x, w.y);
    }
    // ---- propose: x' = x + delta  (or + L delta: rf.py:346-348) -------------
    double xp[D];
    if (a.has_L) {
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) v = fma(m.L[i * D + j], dl[j], v);
        xp[i] = x[i] + v;
      }
    } else {
#pragma unroll
      for (int j = 0; j < D; ++j) xp[j] = x[j] + dl[j];
    }
    // ---- evaluate target ----------------------------------------------------
    const double lpp = mvn_logpdf<D>(xp, m);
    // ---- score / threshold / update (sp_utils.py:19-37) ---------------------
    bool acc;
    double s = nan("");
    double linp = 0.0;
    if (ref_mode) {
      linp = a.log_pscale ? pbx_exp_logp(lpp) : exp(lpp);
      if (gstep == 0) {
        acc = true;
      } else {
        s = fmin(1.0, linp / fmax(PBX_TINY, lin));
        acc = (s >= t);
      }
    } else {
      if (gstep == 0) {
        acc = true;
      } else {
        const double d = lpp - lp;
        acc = (d >= log(t));
        if (a.out_score) s = fmin(1.0, exp(fmin(d, 0.0)));
      }
    }
    if (acc) {
#pragma unroll
      for (int j = 0; j < D; ++j) x[j] = xp[j];
      lp = lpp;
      lin = linp;
      ++nacc;
    }
    // ---- record -------------------------------------------------------------
#pragma unroll
    for (int j = 0; j < D; ++j) {
      ssum[j] += x[j];
      ssq[j] = fma(x[j], x[j], ssq[j]);
    }
    if (a.out_accept) a.out_accept[(int64_t)k * C + c] = acc ? 1 : 0;
    if (a.out_score) a.out_score[(int64_t)k * C + c] = s;
    if ((k + 1) % a.thin == 0) {
      const int64_t r = (k + 1) / a.thin - 1;
      if (a.out_x) {
#pragma unroll
        for (int j = 0; j < D; ++j) a.out_x[(r * D + j) * C + c] = x[j];
      }
      if (a.out_prob) {
        double pv = a.log_pscale ? lp : (ref_mode ? lin : exp(lp));
        a.out_prob[r * C + c] = pv;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < D; ++j) {
    a.state[j * C + c] = x[j];
    if (a.stat_sum) a.stat_sum[j * C + c] += ssum[j];
    if (a.stat_sumsq) a.stat_sumsq[j * C + c] += ssq[j];
  }
  a.state_lp[c] = lp;
  if (a.accept_count) a.accept_count[c] += nacc;
}

template <int D>
static int launch_mh_mvn(pbx_ctx* ctx, const MhMvnArgs& a, const MhMvnConst& m) {
  const int warps = (a.C + 31) / 32;
  // few chains: one warp per CTA so the warps spread over all SMs
  const int block = (warps <= ctx->sm_count * 8) ? 32 : 128;
  const int grid = (a.C + block - 1) / block;
  if (a.inj_delta)
    mh_mvn_kernel<D, true><<<grid, block, 0, ctx->stream>>>(a, m);
  else
    mh_mvn_kernel<D, false><<<grid, block, 0, ctx->stream>>>(a, m);
  PBX_LAUNCH_CHECK(ctx);
  return PBX_OK;
}

static int validate(const pbx_mh_mvn_params* p, const char* who) {
  PBX_REQUIRE(p != nullptr, "%s: null params", who);
  PBX_REQUIRE(p->n_chains >= 1, "%s: n_chains must be >= 1 (got %d)", who, p->n_chains);
  PBX_REQUIRE(p->n_dims >= 1 && p->n_dims <= PBX_MAX_DIMS, "%s: n_dims must be in 1..%d (got %d)",
              who, PBX_MAX_DIMS, p->n_dims);
  PBX_REQUIRE(p->n_steps >= 0, "%s: n_steps must be >= 0", who);
  PBX_REQUIRE(p->thin >= 1, "%s: thin must be >= 1", who);
  PBX_REQUIRE(p->step0 >= 0 && p->chain0 >= 0, "%s: step0/chain0 must be >= 0", who);
  PBX_REQUIRE(p->accept_mode == PBX_ACCEPT_REFERENCE || p->accept_mode == PBX_ACCEPT_LOG,
              "%s: unknown accept_mode %d", who, p->accept_mode);
  PBX_REQUIRE(p->prop_kind == PBX_PROP_NORMAL || p->prop_kind == PBX_PROP_UNIFORM,
              "%s: unknown prop_kind %d", who, p->prop_kind);
  PBX_REQUIRE(p->state && p->state_lp, "%s: state/state_lp are mandatory", who);
  PBX_REQUIRE((p->inj_delta == nullptr) == (p->inj_thresh == nullptr),
              "%s: inj_delta and inj_thresh must be given together", who);
  return PBX_OK;
}

static void fill_const(const pbx_mh_mvn_params* p, MhMvnConst& m) {
  const int D = p->n_dims;
  for (int j = 0; j < PBX_MAX_DIMS; ++j) {
    m.mean[j] = j < D ? p->mean[j] : 0.0;
    m.scale[j] = j < D ? p->prop_scale[j] : 0.0;
  }
  for (int i = 0; i < PBX_MAX_DIMS * PBX_MAX_DIMS; ++i) {
    m.W[i] = i < D * D ? p->whiten[i] : 0.0;
    m.L[i] = i < D * D ? p->prop_mat[i] : 0.0;
  }
  m.norm_c = p->norm_c;
}

static int run_device(pbx_ctx* ctx, const pbx_mh_mvn_params* p) {
  MhMvnConst m;
  fill_const(p, m);
  MhMvnArgs a;
  a.C = p->n_chains; a.T = p->n_steps; a.thin = p->thin;
  a.step0 = p->step0; a.chain0 = p->chain0; a.seed = p->seed;
  a.log_pscale = p->log_pscale; a.accept_mode = p->accept_mode;
  a.prop_kind = p->prop_kind; a.has_L = p->has_prop_mat;
  a.state = p->state; a.state_lp = p->state_lp;
  a.inj_delta = p->inj_delta; a.inj_thresh = p->inj_thresh;
  a.out_x = p->out_x; a.out_prob = p->out_prob;
  a.out_accept = p->out_accept; a.out_score = p->out_score;
  a.accept_count = p->accept_count; a.stat_sum = p->stat_sum; a.stat_sumsq = p->stat_sumsq;
  if (a.T == 0) return PBX_OK;
  switch (p->n_dims) {
    case 1: return launch_mh_mvn<1>(ctx, a, m);
    case 2: return launch_mh_mvn<2>(ctx, a, m);
    case 3: return launch_mh_mvn<3>(ctx, a, m);
    case 4: return launch_mh_mvn<4>(ctx, a, m);
    case 5: return launch_mh_mvn<5>(ctx, a, m);
    case 6: return launch_mh_mvn<6>(ctx, a, m);
    case 7: return launch_mh_mvn<7>(ctx, a, m);
    case 8: return launch_mh_mvn<8>(ctx, a, m);
  }
  pbx_set_error("pbx_mh_mvn_run: unsupported n_dims %d", p->n_dims);
  return PBX_ERR_UNSUPPORTED;
}

extern "C" int pbx_mh_mvn_run(pbx_ctx* ctx, const pbx_mh_mvn_params* p) {
  PBX_REQUIRE(ctx != nullptr, "pbx_mh_mvn_run: null ctx");
  int rc = validate(p, "pbx_mh_mvn_run");
  if (rc) return rc;
  PBX_CUDA(cudaSetDevice(ctx->device));
  PBX_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  rc = run_device(ctx, p);
  if (rc) return rc;
  PBX_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  return PBX_OK;
}

// Whole walk with host buffers: chunks of steps, double-buffered device staging,
// D2H on the copy stream overlapped with the next chunk's kernel.
extern "C" int pbx_mh_mvn_walk_host(pbx_ctx* ctx, const pbx_mh_mvn_params* p,
                                    int32_t chunk_steps) {
  PBX_REQUIRE(ctx != nullptr, "pbx_mh_mvn_walk_host: null ctx");
  int rc = validate(p, "pbx_mh_mvn_walk_host");
  if (rc) return rc;
  PBX_REQUIRE(!p->inj_delta && !p->out_accept && !p->out_score,
              "pbx_mh_mvn_walk_host: injected streams / per-step accept+score outputs are "
              "device-only (use pbx_mh_mvn_run)");
  PBX_CUDA(cudaSetDevice(ctx->device));
  const int64_t C = p->n_chains, D = p->n_dims, T = p->n_steps;
  const int thin = p->thin;
  if (chunk_steps <= 0) chunk_steps = 1024;
  chunk_steps = (chunk_steps + thin - 1) / thin * thin;   // chunk boundaries on record boundaries
  const int64_t rec_per_chunk = chunk_steps / thin;
  const size_t xbytes = p->out_x ? (size_t)rec_per_chunk * D * C * 8 : 0;
  const size_t pbytes = p->out_prob ? (size_t)rec_per_chunk * C * 8 : 0;
  const size_t sbytes = (size_t)D * C * 8;
  // workspace: state, lp, acc, sum, sumsq, 2 x (x chunk, prob chunk)
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = off; off += (b + 255) / 256 * 256; return o; };
  const size_t o_state = take(sbytes), o_lp = take(C * 8), o_acc = take(C * 8),
               o_sum = take(sbytes), o_sq = take(sbytes);
  size_t o_x[2], o_p[2];
  for (int b = 0; b < 2; ++b) { o_x[b] = take(xbytes); o_p[b] = take(pbytes); }
  rc = pbx_ws_reserve(ctx, off);
  if (rc) return rc;
  char* ws = (char*)ctx->ws;
  cudaStream_t st = ctx->stream, cs = ctx->copy_stream;
  PBX_CUDA(cudaMemcpyAsync(ws + o_state, p->state, sbytes, cudaMemcpyHostToDevice, st));
  if (p->step0 > 0)
    PBX_CUDA(cudaMemcpyAsync(ws + o_lp, p->state_lp, C * 8, cudaMemcpyHostToDevice, st));
  PBX_CUDA(cudaMemsetAsync(ws + o_acc, 0, C * 8, st));
  PBX_CUDA(cudaMemsetAsync(ws + o_sum, 0, sbytes, st));
  PBX_CUDA(cudaMemsetAsync(ws + o_sq, 0, sbytes, st));

  cudaEvent_t done[2], copied[2];
  for (int b = 0; b < 2; ++b) {
    PBX_CUDA(cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming));
    PBX_CUDA(cudaEventCreateWithFlags(&copied[b], cudaEventDisableTiming));
  }
  PBX_CUDA(cudaEventRecord(ctx->ev0, st));
  int64_t rec_done = 0;
  int chunk = 0;
  for (int64_t t0 = 0; t0 < T; t0 += chunk_steps, ++chunk) {
    const int b = chunk & 1;
    const int64_t nt = (T - t0 < chunk_steps) ? (T - t0) : chunk_steps;
    const int64_t nrec = nt / thin;
    if (chunk >= 2) PBX_CUDA(cudaStreamWaitEvent(st, copied[b], 0));
    pbx_mh_mvn_params q = *p;
    q.n_steps = (int32_t)nt;
    q.step0 = p->step0 + t0;
    q.state = (double*)(ws + o_state);
    q.state_lp = (double*)(ws + o_lp);
    q.out_x = p->out_x ? (double*)(ws + o_x[b]) : nullptr;
    q.out_prob = p->out_prob ? (double*)(ws + o_p[b]) : nullptr;
    q.accept_count = (int64_t*)(ws + o_acc);
    q.stat_sum = (double*)(ws + o_sum);
    q.stat_sumsq = (double*)(ws + o_sq);
    rc = run_device(ctx, &q);
    if (rc) return rc;
    PBX_CUDA(cudaEventRecord(done[b], st));
    PBX_CUDA(cudaStreamWaitEvent(cs, done[b], 0));
    if (p->out_x && nrec)
      PBX_CUDA(cudaMemcpyAsync(p->out_x + rec_done * D * C, ws + o_x[b], (size_t)nrec * D * C * 8,
                               cudaMemcpyDeviceToHost, cs));
    if (p->out_prob && nrec)
      PBX_CUDA(cudaMemcpyAsync(p->out_prob + rec_done * C, ws + o_p[b], (size_t)nrec * C * 8,
                               cudaMemcpyDeviceToHost, cs));
    PBX_CUDA(cudaEventRecord(copied[b], cs));
    rec_done += nrec;
  }
  PBX_CUDA(cudaEventRecord(ctx->ev1, st));
  PBX_CUDA(cudaMemcpyAsync(p->state, ws + o_state, sbytes, cudaMemcpyDeviceToHost, st));
  PBX_CUDA(cudaMemcpyAsync(p->state_lp, ws + o_lp, C * 8, cudaMemcpyDeviceToHost, st));
  if (p->accept_count)
    PBX_CUDA(cudaMemcpyAsync(p->accept_count, ws + o_acc, C * 8, cudaMemcpyDeviceToHost, st));
  if (p->stat_sum)
    PBX_CUDA(cudaMemcpyAsync(p->stat_sum, ws + o_sum, sbytes, cudaMemcpyDeviceToHost, st));
  if (p->stat_sumsq)
    PBX_CUDA(cudaMemcpyAsync(p->stat_sumsq, ws + o_sq, sbytes, cudaMemcpyDeviceToHost, st));
  PBX_CUDA(cudaStreamSynchronize(st));
  PBX_CUDA(cudaStreamSynchronize(cs));
  for (int b = 0; b < 2; ++b) {
    cudaEventDestroy(done[b]);
    cudaEventDestroy(copied[b]);
  }
  return PBX_OK;
}
