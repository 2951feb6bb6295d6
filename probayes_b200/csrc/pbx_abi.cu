// pbx_abi.cu -- context, error reporting, device info, small utility kernels.
#include <stdarg.h>
#include <string.h>
#include "pbx_common.cuh"

static thread_local char g_err[1024] = "";

void pbx_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int pbx_ws_reserve(pbx_ctx* ctx, size_t bytes) {
  if (ctx->ws_bytes >= bytes) return PBX_OK;
  if (ctx->ws) {
    PBX_CUDA(cudaStreamSynchronize(ctx->stream));
    PBX_CUDA(cudaFree(ctx->ws));
    ctx->ws = nullptr;
    ctx->ws_bytes = 0;
  }
  size_t want = bytes < (1u << 20) ? (1u << 20) : bytes;
  cudaError_t e = cudaMalloc(&ctx->ws, want);
  if (e != cudaSuccess) {
    pbx_set_error("workspace cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    return PBX_ERR_NOMEM;
  }
  ctx->ws_bytes = want;
  return PBX_OK;
}

extern "C" {

int pbx_version(void) { return PBX_VERSION; }

const char* pbx_last_error(void) { return g_err; }

int pbx_device_count(int* n) {
  PBX_REQUIRE(n != nullptr, "pbx_device_count: null out");
  cudaError_t e = cudaGetDeviceCount(n);
  if (e != cudaSuccess) {
    *n = 0;
    pbx_set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return PBX_ERR_CUDA;
  }
  return PBX_OK;
}

int pbx_device_info(int device, pbx_devinfo* out) {
  PBX_REQUIRE(out != nullptr, "pbx_device_info: null out");
  cudaDeviceProp prop;
  PBX_CUDA(cudaGetDeviceProperties(&prop, device));
  memset(out, 0, sizeof(*out));
  out->device = device;
  out->cc_major = prop.major;
  out->cc_minor = prop.minor;
  out->sm_count = prop.multiProcessorCount;
  out->l2_bytes_mb = (int32_t)(prop.l2CacheSize >> 20);
  out->smem_per_block_optin = (int32_t)prop.sharedMemPerBlockOptin;
  out->global_mem_bytes = (int64_t)prop.totalGlobalMem;
  strncpy(out->name, prop.name, sizeof(out->name) - 1);
  return PBX_OK;
}

int pbx_ctx_create(int device, void* stream, int32_t own_stream, pbx_ctx** out) {
  PBX_REQUIRE(out != nullptr, "pbx_ctx_create: null out");
  *out = nullptr;
  int n = 0;
  PBX_CUDA(cudaGetDeviceCount(&n));
  PBX_REQUIRE(device >= 0 && device < n, "pbx_ctx_create: device %d out of range (%d present)",
              device, n);
  PBX_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  PBX_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    pbx_set_error("libpbx is built for sm_100a only; device %d is sm_%d%d (%s)", device,
                  prop.major, prop.minor, prop.name);
    return PBX_ERR_UNSUPPORTED;
  }
  pbx_ctx* c = new pbx_ctx();
  memset(c, 0, sizeof(*c));
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  cudaError_t e = cudaSuccess;
  if (own_stream) {
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    c->own_stream = (e == cudaSuccess);
  } else {
    c->stream = (cudaStream_t)stream;                  // NULL = the default stream
    c->own_stream = false;
  }
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
  if (e != cudaSuccess) {                              // no half-built context is leaked
    pbx_set_error("pbx_ctx_create: %s", cudaGetErrorString(e));
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return PBX_ERR_CUDA;
  }
  *out = c;
  return PBX_OK;
}

int pbx_ctx_destroy(pbx_ctx* ctx) {
  if (!ctx) return PBX_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaStreamSynchronize(ctx->copy_stream);
  if (ctx->ws) cudaFree(ctx->ws);
  if (ctx->ticket) cudaFree(ctx->ticket);
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  cudaStreamDestroy(ctx->copy_stream);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return PBX_OK;
}

int pbx_ctx_sync(pbx_ctx* ctx) {
  PBX_REQUIRE(ctx != nullptr, "pbx_ctx_sync: null ctx");
  PBX_CUDA(cudaStreamSynchronize(ctx->stream));
  return PBX_OK;
}

int64_t pbx_ctx_launch_count(pbx_ctx* ctx) { return ctx ? ctx->launches : -1; }

int pbx_ctx_last_kernel_ms(pbx_ctx* ctx, float* ms) {
  PBX_REQUIRE(ctx != nullptr && ms != nullptr, "pbx_ctx_last_kernel_ms: null argument");
  PBX_CUDA(cudaEventSynchronize(ctx->ev1));
  PBX_CUDA(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
  return PBX_OK;
}

int pbx_host_alloc(size_t bytes, void** out) {
  PBX_REQUIRE(out != nullptr, "pbx_host_alloc: null out");
  PBX_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
  return PBX_OK;
}

int pbx_host_free(void* p) {
  if (p) PBX_CUDA(cudaFreeHost(p));
  return PBX_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// elementwise clamped log / exp (PD.rescaled: probayes/pd.py:496-499)
// ---------------------------------------------------------------------------
template <bool kLog>
__global__ void __launch_bounds__(256) pbx_rescale_kernel(double* __restrict__ v, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) v[i] = kLog ? pbx_log_prob(v[i]) : pbx_exp_logp(v[i]);
}

// ---------------------------------------------------------------------------
// chain summaries: one block per dim, fixed-order tree => deterministic.
// The within-chain variance is formed from the raw running sums the walk kernels keep,
// (sum x^2 - sum x * mean) / (T - 1): relative error ~ eps * mean^2 / var, i.e. fine for
// posteriors whose spread is not many orders below their location (1e-8 relative at
// |mean| / sd = 1e4); documented limitation.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pbx_chain_stats_kernel(const double* __restrict__ ssum,
                                                              const double* __restrict__ ssq,
                                                              int64_t C, double T,
                                                              double* __restrict__ out) {
  int j = blockIdx.x;
  double a0 = 0, a1 = 0, a2 = 0;
  for (int64_t c = threadIdx.x; c < C; c += blockDim.x) {
    double s = ssum[(int64_t)j * C + c], q = ssq[(int64_t)j * C + c];
    double m = s / T;
    double v = (T > 1.0) ? (q - s * m) / (T - 1.0) : 0.0;
    a0 += m; a1 += m * m; a2 += v;
  }
  __shared__ double sh[3][256];
  sh[0][threadIdx.x] = a0; sh[1][threadIdx.x] = a1; sh[2][threadIdx.x] = a2;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
      sh[2][threadIdx.x] += sh[2][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[j * 4 + 0] = sh[0][0]; out[j * 4 + 1] = sh[1][0]; out[j * 4 + 2] = sh[2][0];
    out[j * 4 + 3] = (double)C;
  }
}

// ---------------------------------------------------------------------------
// FP64 FMA peak: 8 independent dependent-FMA chains per thread
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pbx_fp64_peak_kernel(double* out, int iters, double a,
                                                            double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5,
         x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 12345.678) out[0] = s;   // keep the chain alive
}

// dependent chain of n DFMA / DADD in one warp, timed with clock64
__global__ void pbx_fp64_latency_kernel(double* out, int n, double a, double b) {
  double x = (double)threadIdx.x * 1e-3, y = x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) x = fma(x, a, b);
  long long t1 = clock64();
  for (int i = 0; i < n; ++i) y = y + b;
  long long t2 = clock64();
  if (threadIdx.x == 0) {
    out[0] = (double)(t1 - t0) / n;
    out[1] = (double)(t2 - t1) / n;
    out[2] = x + y;
  }
}

extern "C" {

int pbx_fp64_dep_latency(pbx_ctx* ctx, double* dfma_cycles, double* dadd_cycles) {
  PBX_REQUIRE(ctx && dfma_cycles && dadd_cycles, "pbx_fp64_dep_latency: null argument");
  PBX_CUDA(cudaSetDevice(ctx->device));
  {
    int rc = pbx_ws_reserve(ctx, 64);
    if (rc) return rc;
  }
  double h[3];
  for (int rep = 0; rep < 2; ++rep) {
    pbx_fp64_latency_kernel<<<1, 32, 0, ctx->stream>>>((double*)ctx->ws, 4096, 1.0000001, 1e-9);
    PBX_LAUNCH_CHECK(ctx);
  }
  PBX_CUDA(cudaMemcpyAsync(h, ctx->ws, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  PBX_CUDA(cudaStreamSynchronize(ctx->stream));
  *dfma_cycles = h[0];
  *dadd_cycles = h[1];
  return PBX_OK;
}

int pbx_log_prob_inplace(pbx_ctx* ctx, double* v, int64_t n) {
  PBX_REQUIRE(ctx && v && n >= 0, "pbx_log_prob_inplace: bad argument");
  if (n == 0) return PBX_OK;
  PBX_CUDA(cudaSetDevice(ctx->device));
  int grid = (int)((n + 255) / 256);
  if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
  pbx_rescale_kernel<true><<<grid, 256, 0, ctx->stream>>>(v, n);
  PBX_LAUNCH_CHECK(ctx);
  return PBX_OK;
}

int pbx_exp_logp_inplace(pbx_ctx* ctx, double* v, int64_t n) {
  PBX_REQUIRE(ctx && v && n >= 0, "pbx_exp_logp_inplace: bad argument");
  if (n == 0) return PBX_OK;
  PBX_CUDA(cudaSetDevice(ctx->device));
  int grid = (int)((n + 255) / 256);
  if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
  pbx_rescale_kernel<false><<<grid, 256, 0, ctx->stream>>>(v, n);
  PBX_LAUNCH_CHECK(ctx);
  return PBX_OK;
}

int pbx_reduce_chain_stats(pbx_ctx* ctx, const double* stat_sum, const double* stat_sumsq,
                           int32_t n_dims, int64_t n_chains, int64_t n_steps, double* out) {
  PBX_REQUIRE(ctx && stat_sum && stat_sumsq && out, "pbx_reduce_chain_stats: null argument");
  PBX_REQUIRE(n_dims >= 1 && n_chains >= 1 && n_steps >= 1,
              "pbx_reduce_chain_stats: sizes must be positive");
  PBX_CUDA(cudaSetDevice(ctx->device));
  pbx_chain_stats_kernel<<<n_dims, 256, 0, ctx->stream>>>(stat_sum, stat_sumsq, n_chains,
                                                         (double)n_steps, out);
  PBX_LAUNCH_CHECK(ctx);
  return PBX_OK;
}

int pbx_fp64_peak(pbx_ctx* ctx, double* tflops) {
  PBX_REQUIRE(ctx && tflops, "pbx_fp64_peak: null argument");
  PBX_CUDA(cudaSetDevice(ctx->device));
  {
    int rc = pbx_ws_reserve(ctx, 64);
    if (rc) return rc;
  }
  const int iters = 4096, blocks = ctx->sm_count * 8, threads = 256;
  double best = 0;
  for (int rep = 0; rep < 4; ++rep) {
    PBX_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    pbx_fp64_peak_kernel<<<blocks, threads, 0, ctx->stream>>>((double*)ctx->ws, iters,
                                                              1.0000001, 1e-9);
    PBX_LAUNCH_CHECK(ctx);
    PBX_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    PBX_CUDA(cudaEventSynchronize(ctx->ev1));
    float ms = 0;
    PBX_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
    double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  *tflops = best;
  return PBX_OK;
}

}  // extern "C"
