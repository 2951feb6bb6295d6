// pbx_common.cuh -- shared host/device helpers for libpbx (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "pbx.h"

// ---------------------------------------------------------------------------
// context + error plumbing (host)
// ---------------------------------------------------------------------------
struct pbx_ctx {
  int device;
  cudaStream_t stream;       // compute stream (borrowed or owned)
  cudaStream_t copy_stream;  // owned; D2H of the *_host entry points
  bool own_stream;
  int sm_count;
  int64_t launches;
  cudaEvent_t ev0, ev1;      // bracket the kernels of the most recent *_run call
  void* ws;                  // device workspace
  size_t ws_bytes;
};

void pbx_set_error(const char* fmt, ...);
int pbx_ws_reserve(pbx_ctx* ctx, size_t bytes);

#define PBX_CUDA(call)                                                          \
  do {                                                                          \
    cudaError_t _e = (call);                                                    \
    if (_e != cudaSuccess) {                                                    \
      pbx_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call,          \
                    cudaGetErrorString(_e));                                    \
      return PBX_ERR_CUDA;                                                      \
    }                                                                           \
  } while (0)

#define PBX_REQUIRE(cond, ...)                                                  \
  do {                                                                          \
    if (!(cond)) {                                                              \
      pbx_set_error(__VA_ARGS__);                                               \
      return PBX_ERR_INVALID;                                                   \
    }                                                                           \
  } while (0)

#define PBX_LAUNCH_CHECK(ctx)                                                   \
  do {                                                                          \
    (ctx)->launches++;                                                          \
    PBX_CUDA(cudaGetLastError());                                               \
  } while (0)

// ---------------------------------------------------------------------------
// probayes/constants.py:9-32 (fp64)
// ---------------------------------------------------------------------------
#define PBX_TINY 2.2250738585072014e-308
#define PBX_HUGE 1.7976931348623158e+308
#define PBX_LOG_HUGE 709.782712893384   /* log(PBX_HUGE) */
#define PBX_LOG_SQRT_2PI 0.91893853320467274178

// clamped log / exp of probayes/pscales.py:44-65 (NaN -> +huge, as the reference)
__device__ __forceinline__ double pbx_log_prob(double p) {
  return (p >= PBX_TINY) ? log(p) : -PBX_HUGE;
}
__device__ __forceinline__ double pbx_exp_logp(double l) {
  return (l <= PBX_LOG_HUGE) ? exp(l) : PBX_HUGE;
}

// ---------------------------------------------------------------------------
// Philox4x32-10 counter RNG.  Stream layout documented in oracle/philox.py:
//   key = (seed lo, seed hi); counter = (step lo, step hi, chain, slot)
// ---------------------------------------------------------------------------
#define PBX_SLOT_THRESH 255u

struct pbx_u4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ pbx_u4 pbx_philox(uint32_t c0, uint32_t c1, uint32_t c2,
                                                      uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  pbx_u4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
  return o;
}

__host__ __device__ __forceinline__ pbx_u4 pbx_block(uint64_t seed, uint64_t step, uint32_t chain,
                                                     uint32_t slot) {
  return pbx_philox((uint32_t)step, (uint32_t)(step >> 32), chain, slot, (uint32_t)seed,
                    (uint32_t)(seed >> 32));
}

// (a, b) -> (2k+1) * 2^-53 with 52-bit k: exact in fp64, strictly inside (0,1)
__host__ __device__ __forceinline__ double pbx_u01(uint32_t a, uint32_t b) {
  uint64_t k = ((uint64_t)a << 20) | (uint64_t)(b >> 12);
  return (double)(2 * k + 1) * 1.1102230246251565e-16;   // 2^-53
}

#ifdef __CUDACC__
// Box-Muller pair from one Philox block
__device__ __forceinline__ void pbx_normal_pair(pbx_u4 w, double& z0, double& z1) {
  double u1 = pbx_u01(w.x, w.y), u2 = pbx_u01(w.z, w.w);
  double r = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  z0 = r * c;
  z1 = r * s;
}
#endif
