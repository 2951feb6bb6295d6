// pbx_post.cu -- K6: PD post-processing on large grids / sample sets (sm_100a), and the
// box sampler of ordinary Monte Carlo random sampling.
//
//   pbx_argsort_f64      PD.sorted's np.argsort           (probayes/pd.py:464-493)
//   pbx_gather_f64 / pbx_take_axis_f64   its fancy-indexing re-orderings (pd.py:476-492)
//   pbx_cumprob_f64      PD.quantile's rescale + cumsum + div_prob   (pd.py:426-429)
//   pbx_digitize_f64     np.maximum(0, np.digitize(q, cum) - 1)      (pd.py:430)
//   pbx_expectation_f64  PD.expectation's rescale + sum(prob * val)  (pd.py:387-402)
//   pbx_box_sample       Variable.evaluate({0})   (variable.py:558-583, vtypes.py:186)
//
// All of it is HBM-bound streaming work: coalesced 8/16-byte accesses, shared-memory
// staging where the access pattern is data dependent (the radix scatter), grids sized
// from the tile count, every reduction / scan in a fixed order (bit-reproducible).
#include <algorithm>
#include <cmath>
#include "pbx_common.cuh"

// ===========================================================================
// Radix argsort: LSD, 8-bit digits, stable.  Per executed pass:
//   rs_hist     tile digit counts            -> counts[digit][tile]   (read 8 B/key)
//   rs_rowscan  exclusive scan of each digit row over the tiles (in place)
//   rs_scatter  stable ranks (warp match + per-warp counters), tile re-ordered by
//               digit in shared memory, then written out in coalesced runs
//               (read 12 B/key, write 12 B/key)
// A digit position whose value is the same for every key (sign/exponent bytes of
// same-magnitude data) is skipped: its pass would be the identity.  Which passes
// run, and which buffer each reads/writes, is decided ON THE DEVICE (rs_plan) from
// the bitwise AND / OR of all keys (one streaming read), so the call stays
// asynchronous -- skipped passes are launches that return at once.
// Keys of equal digit inside a warp are matched with 8 ballots (one per digit bit):
// the hardware match.any instruction is several times slower than that.
// ===========================================================================
#define RS_THREADS 256
#ifndef RS_ITEMS
#define RS_ITEMS 16
#endif
#define RS_TILE (RS_THREADS * RS_ITEMS)
#define RS_WARPS (RS_THREADS / 32)
#define RS_PASSES 8
#ifndef RS_MINB
#define RS_MINB 3
#endif

struct RsPlan {
  int skip[RS_PASSES];
  int src[RS_PASSES];        // 0 = caller's keys (index implicit), 1 / 2 = workspace A / B
  int dst[RS_PASSES];        // 1 / 2 = workspace A / B, 3 = final (order, keys_sorted)
  int n_exec;
  unsigned dtot[256];               // digit totals of the pass in flight (rs_rowscan)
};

struct RsBufs {
  const unsigned long long* keys;   // caller's doubles, read as bits
  unsigned long long* ka;
  unsigned long long* kb;
  int* ia;
  int* ib;
  int* order;
  unsigned long long* keys_sorted;  // may be null
};

// order-preserving map double bits -> unsigned
__device__ __forceinline__ unsigned long long rs_encode(unsigned long long b) {
  return b ^ ((b >> 63) ? 0xFFFFFFFFFFFFFFFFull : 0x8000000000000000ull);
}
__device__ __forceinline__ unsigned long long rs_decode(unsigned long long u) {
  return u ^ ((u >> 63) ? 0x8000000000000000ull : 0xFFFFFFFFFFFFFFFFull);
}

// lanes of the warp holding the same 8-bit digit (all 32 lanes must call)
__device__ __forceinline__ unsigned rs_match8(unsigned d) {
  unsigned m = 0xffffffffu;
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const bool bit = (d >> b) & 1u;
    const unsigned bal = __ballot_sync(0xffffffffu, bit);
    m &= bit ? bal : ~bal;
  }
  return m;
}

// bitwise AND and OR of all encoded keys: byte p of (and ^ or) is zero <=> digit p constant
__global__ void __launch_bounds__(256) rs_andor_kernel(const unsigned long long* __restrict__ keys,
                                                       long long n,
                                                       unsigned long long* __restrict__ andor) {
  unsigned long long a = ~0ull, o = 0ull;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long u = rs_encode(keys[i]);
    a &= u;
    o |= u;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    a &= __shfl_xor_sync(0xffffffffu, a, off);
    o |= __shfl_xor_sync(0xffffffffu, o, off);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAnd(&andor[0], a);
    atomicOr(&andor[1], o);
  }
}

__global__ void rs_plan_kernel(const unsigned long long* __restrict__ andor, RsPlan* plan) {
  if (threadIdx.x != 0) return;
  const unsigned long long diff = andor[0] ^ andor[1];
  int k = 0, last = -1;
  for (int p = 0; p < RS_PASSES; ++p)
    if ((diff >> (8 * p)) & 255ull) last = p;
  for (int p = 0; p < RS_PASSES; ++p) {
    const int skip = ((diff >> (8 * p)) & 255ull) == 0;
    plan->skip[p] = skip;
    if (skip) continue;
    plan->src[p] = (k == 0) ? 0 : 1 + ((k - 1) & 1);
    plan->dst[p] = (p == last) ? 3 : 1 + (k & 1);
    ++k;
  }
  plan->n_exec = k;
}

__device__ __forceinline__ unsigned long long rs_load_key(const RsBufs& b, int src, long long e) {
  if (src == 0) return rs_encode(b.keys[e]);
  return (src == 1 ? b.ka : b.kb)[e];
}

// Tile digit counts without votes or contended atomics: every LANE owns a private copy
// of the histogram, laid out [digit pair][lane] so that lane l only ever touches bank l
// (two 16-bit counters per word; a lane sees at most 16 x 8 = 128 keys of a tile).  The
// shared-memory atomics of a warp therefore never collide, whatever the digit skew.
__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const RsBufs b, long long n, int pass,
                                                             const RsPlan* __restrict__ plan,
                                                             unsigned* __restrict__ counts,
                                                             int ntiles) {
  if (plan->skip[pass]) return;
  __shared__ unsigned h[128 * 32];
#pragma unroll
  for (int i = 0; i < 16; ++i) h[i * RS_THREADS + threadIdx.x] = 0;
  const int src = plan->src[pass];
  const int lane = threadIdx.x & 31;
  const int shift = 8 * pass;
  const long long base = (long long)blockIdx.x * RS_TILE;
  const unsigned long long* kp = (src == 0) ? b.keys : (src == 1 ? b.ka : b.kb);
  unsigned long long k[RS_ITEMS];
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {             // all loads in flight at once
    const long long e = base + r * RS_THREADS + threadIdx.x;
    k[r] = (e < n) ? kp[e] : 0ull;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const long long e = base + r * RS_THREADS + threadIdx.x;
    const unsigned long long u = (src == 0) ? rs_encode(k[r]) : k[r];
    const unsigned d = (unsigned)(u >> shift) & 255u;
    if (e < n) atomicAdd(&h[(d >> 1) * 32 + lane], 1u << ((d & 1u) * 16));
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    unsigned acc = 0;                              // both halves stay below 2^16 (<= 4096)
#pragma unroll
    for (int j = 0; j < 32; ++j) acc += h[threadIdx.x * 32 + ((j + threadIdx.x) & 31)];
    counts[(size_t)(2 * threadIdx.x) * ntiles + blockIdx.x] = acc & 0xFFFFu;
    counts[(size_t)(2 * threadIdx.x + 1) * ntiles + blockIdx.x] = acc >> 16;
  }
}

// CTA d: exclusive scan of counts[d][0..ntiles) in place
__global__ void __launch_bounds__(256) rs_rowscan_kernel(unsigned* __restrict__ counts, int ntiles,
                                                         int pass, RsPlan* __restrict__ plan) {
  if (plan->skip[pass]) return;
  unsigned* plan_dtot = plan->dtot;
  __shared__ unsigned s_w[8];
  __shared__ unsigned s_carry;
  unsigned* row = counts + (size_t)blockIdx.x * ntiles;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int t0 = 0; t0 < ntiles; t0 += 256 * 4) {
    unsigned v[4], sum = 0;
    const int i0 = t0 + threadIdx.x * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[k] = (i0 + k < ntiles) ? row[i0 + k] : 0u;
      sum += v[k];
    }
    unsigned inc = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, inc, off);
      if (lane >= off) inc += t;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    unsigned wbase = s_carry;
    for (int w = 0; w < warp; ++w) wbase += s_w[w];
    unsigned run = wbase + inc - sum;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i0 + k < ntiles) row[i0 + k] = run;
      run += v[k];
    }
    __syncthreads();
    if (threadIdx.x == 255) s_carry = run;
    __syncthreads();
  }
  if (threadIdx.x == 0) plan_dtot[blockIdx.x] = s_carry;
}

struct RsSmem {
  unsigned long long keys[RS_TILE];   // tile re-ordered by digit; re-used for the indices
  unsigned wc[RS_WARPS][256];         // per-warp digit counters -> local offsets
  unsigned gofs[256];                 // global position of local slot j of digit d = gofs[d] + j
  unsigned wsum[RS_WARPS];
  unsigned gsum[RS_WARPS];
};

__global__ void __launch_bounds__(RS_THREADS, RS_MINB) rs_scatter_kernel(const RsBufs b, long long n,
                                                                   int pass,
                                                                   const RsPlan* __restrict__ plan,
                                                                   const unsigned* __restrict__ counts,
                                                                   int ntiles) {
  if (plan->skip[pass]) return;
  extern __shared__ __align__(16) unsigned char rs_raw[];
  RsSmem& s = *reinterpret_cast<RsSmem*>(rs_raw);
  const int src = plan->src[pass], dst = plan->dst[pass];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int shift = 8 * pass;
  const long long base = (long long)blockIdx.x * RS_TILE;
  const int tile_n = (int)min((long long)RS_TILE, n - base);
  // thread = digit: this tile's exclusive count and the digit's global total (loaded
  // early so that their latency hides behind the key loads)
  const unsigned tile_ofs = counts[(size_t)tid * ntiles + blockIdx.x];
  const unsigned dt = plan->dtot[tid];

#pragma unroll
  for (int i = 0; i < RS_WARPS; ++i) s.wc[i][tid] = 0;

  // tile order: warp w owns elements [w*512, (w+1)*512), item r is element r*32 + lane of it
  const unsigned long long* kp = (src == 0) ? b.keys : (src == 1 ? b.ka : b.kb);
  unsigned long long key[RS_ITEMS];
  const int seg = warp * (RS_ITEMS * 32);
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const int t = seg + r * 32 + lane;
    key[r] = (t < tile_n) ? kp[base + t] : 0ull;
  }
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const int t = seg + r * 32 + lane;
    if (src == 0) key[r] = rs_encode(key[r]);
    if (t >= tile_n) key[r] = 0xFFFFFFFFFFFFFFFFull;   // padding ranks after every real key
  }
  __syncthreads();

  // stable rank of each key among the keys of its digit within the warp segment: the
  // first lane of each digit group adds the group's size to the warp's running counter
  // with one shared-memory atomic and broadcasts the old value to its group
  unsigned short rank[RS_ITEMS];
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const unsigned d = (unsigned)(key[r] >> shift) & 255u;
    const unsigned m = rs_match8(d);
    const int leader = __ffs(m) - 1;
    unsigned prev = 0;
    if (lane == leader) prev = atomicAdd(&s.wc[warp][d], (unsigned)__popc(m));
    prev = __shfl_sync(0xffffffffu, prev, leader);
    rank[r] = (unsigned short)(prev + __popc(m & lt));
  }
  __syncthreads();

  // thread d: warp counters -> exclusive offsets over the warps; tile count of digit d
  unsigned cnt = 0;
#pragma unroll
  for (int w = 0; w < RS_WARPS; ++w) {
    const unsigned t = s.wc[w][tid];
    s.wc[w][tid] = cnt;
    cnt += t;
  }
  // exclusive scans over the 256 digits: the tile's counts, and the pass's digit totals
  unsigned inc = cnt, ginc = dt;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, inc, off);
    const unsigned g = __shfl_up_sync(0xffffffffu, ginc, off);
    if (lane >= off) {
      inc += t;
      ginc += g;
    }
  }
  if (lane == 31) {
    s.wsum[warp] = inc;
    s.gsum[warp] = ginc;
  }
  __syncthreads();
  unsigned excl = inc - cnt, gbase = ginc - dt;
  for (int w = 0; w < warp; ++w) {
    excl += s.wsum[w];
    gbase += s.gsum[w];
  }
  s.gofs[tid] = gbase + tile_ofs - excl;               // mod 2^32; + j >= excl restores it
#pragma unroll
  for (int w = 0; w < RS_WARPS; ++w) s.wc[w][tid] += excl;
  __syncthreads();

  // re-order the tile's keys by digit in shared memory; rank[] becomes the local slot
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const unsigned d = (unsigned)(key[r] >> shift) & 255u;
    const unsigned lp = s.wc[warp][d] + rank[r];
    rank[r] = (unsigned short)lp;
    s.keys[lp] = key[r];
  }
  // the indices travelling with the keys (their loads overlap the key write-out)
  int idv[RS_ITEMS];
  {
    const int* src_idx = (src == 1) ? b.ia : b.ib;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
      const int t = seg + r * 32 + lane;
      idv[r] = (src == 0) ? (int)(base + t) : ((t < tile_n) ? src_idx[base + t] : 0);
    }
  }
  __syncthreads();

  // write out: consecutive threads -> consecutive slots -> runs of one digit -> coalesced
  unsigned long long* dk = (dst == 1) ? b.ka : (dst == 2 ? b.kb : b.keys_sorted);
  int* di = (dst == 1) ? b.ia : (dst == 2 ? b.ib : b.order);
  const bool final_pass = dst == 3;
  unsigned g[RS_ITEMS];
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) {
    const int j = tid + i * RS_THREADS;
    const unsigned long long k = s.keys[j];
    g[i] = s.gofs[(unsigned)(k >> shift) & 255u] + (unsigned)j;
    if (dk && j < tile_n) dk[g[i]] = final_pass ? rs_decode(k) : k;
  }
  __syncthreads();
  int* sidx = reinterpret_cast<int*>(s.keys);
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) sidx[rank[r]] = idv[r];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) {
    const int j = tid + i * RS_THREADS;
    if (j < tile_n) di[g[i]] = sidx[j];
  }
}

// every digit position constant (all keys equal, or n <= 1): the identity order
__global__ void rs_identity_kernel(const RsBufs b, long long n, const RsPlan* __restrict__ plan) {
  if (plan->n_exec != 0) return;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    b.order[i] = (int)i;
    if (b.keys_sorted) b.keys_sorted[i] = b.keys[i];
  }
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

struct RsLayout {
  size_t ka, kb, ia, ib, counts, ghist, plan, total;
  int ntiles;
};
static RsLayout rs_layout(int64_t n) {
  RsLayout L;
  L.ntiles = (int)((n + RS_TILE - 1) / RS_TILE);
  if (L.ntiles < 1) L.ntiles = 1;
  size_t o = 0;
  L.ka = o; o += align256((size_t)n * 8);
  L.kb = o; o += align256((size_t)n * 8);
  L.ia = o; o += align256((size_t)n * 4);
  L.ib = o; o += align256((size_t)n * 4);
  L.counts = o; o += align256((size_t)256 * L.ntiles * 4);
  L.ghist = o; o += align256(16);                 // AND / OR of the encoded keys
  L.plan = o; o += align256(sizeof(RsPlan));
  L.total = o;
  return L;
}

extern "C" int pbx_argsort_workspace_bytes(int64_t n, size_t* bytes) {
  PBX_REQUIRE(bytes != nullptr, "pbx_argsort_workspace_bytes: null out");
  PBX_REQUIRE(n >= 0 && n < ((int64_t)1 << 31), "pbx_argsort_workspace_bytes: n must be in [0, 2^31)");
  *bytes = rs_layout(n).total;
  return PBX_OK;
}

extern "C" int pbx_argsort_f64(pbx_ctx* ctx, const double* keys, int64_t n, int32_t* order,
                               double* keys_sorted, void* workspace, size_t workspace_bytes) {
  PBX_REQUIRE(ctx != nullptr, "pbx_argsort_f64: null context");
  PBX_REQUIRE(n >= 0 && n < ((int64_t)1 << 31), "pbx_argsort_f64: n must be in [0, 2^31) (got %lld)",
              (long long)n);
  if (n == 0) return PBX_OK;
  PBX_REQUIRE(keys && order, "pbx_argsort_f64: keys and order are mandatory");
  const RsLayout L = rs_layout(n);
  PBX_REQUIRE(workspace != nullptr && workspace_bytes >= L.total,
              "pbx_argsort_f64: workspace of %zu bytes needed (got %zu)", L.total, workspace_bytes);
  PBX_CUDA(cudaSetDevice(ctx->device));
  unsigned char* w = static_cast<unsigned char*>(workspace);
  RsBufs b;
  b.keys = reinterpret_cast<const unsigned long long*>(keys);
  b.ka = reinterpret_cast<unsigned long long*>(w + L.ka);
  b.kb = reinterpret_cast<unsigned long long*>(w + L.kb);
  b.ia = reinterpret_cast<int*>(w + L.ia);
  b.ib = reinterpret_cast<int*>(w + L.ib);
  b.order = order;
  b.keys_sorted = reinterpret_cast<unsigned long long*>(keys_sorted);
  unsigned* counts = reinterpret_cast<unsigned*>(w + L.counts);
  unsigned long long* andor = reinterpret_cast<unsigned long long*>(w + L.ghist);
  RsPlan* plan = reinterpret_cast<RsPlan*>(w + L.plan);

  PBX_CUDA(cudaMemsetAsync(andor, 0xFF, 8, ctx->stream));
  PBX_CUDA(cudaMemsetAsync(andor + 1, 0, 8, ctx->stream));
  const int ao_grid = (int)std::min<int64_t>((n + 256 * 8 - 1) / (256 * 8), (int64_t)ctx->sm_count * 8);
  rs_andor_kernel<<<ao_grid, 256, 0, ctx->stream>>>(b.keys, (long long)n, andor);
  PBX_LAUNCH_CHECK(ctx);
  rs_plan_kernel<<<1, 32, 0, ctx->stream>>>(andor, plan);
  PBX_LAUNCH_CHECK(ctx);
  PBX_CUDA(cudaFuncSetAttribute(rs_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(RsSmem)));
  for (int p = 0; p < RS_PASSES; ++p) {
    rs_hist_kernel<<<L.ntiles, RS_THREADS, 0, ctx->stream>>>(b, (long long)n, p, plan, counts,
                                                            L.ntiles);
    PBX_LAUNCH_CHECK(ctx);
    rs_rowscan_kernel<<<256, 256, 0, ctx->stream>>>(counts, L.ntiles, p, plan);
    PBX_LAUNCH_CHECK(ctx);
    rs_scatter_kernel<<<L.ntiles, RS_THREADS, sizeof(RsSmem), ctx->stream>>>(b, (long long)n, p,
                                                                            plan, counts, L.ntiles);
    PBX_LAUNCH_CHECK(ctx);
  }
  const int id_grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 8);
  rs_identity_kernel<<<id_grid, 256, 0, ctx->stream>>>(b, (long long)n, plan);
  PBX_LAUNCH_CHECK(ctx);
  return PBX_OK;
}

// ===========================================================================
// gather / take along an axis
// ===========================================================================
__global__ void __launch_bounds__(256) gather_kernel(const double* __restrict__ src,
                                                     const int* __restrict__ idx, long long n,
                                                     double* __restrict__ dst) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[idx[i]];
}

// axis 0: whole rows move (coalesced both sides); axis 1: reads are a permutation
// inside each row (the row stays L2/L1-resident while it is consumed)
__global__ void __launch_bounds__(256) take_axis_kernel(const double* __restrict__ src,
                                                        long long rows, long long cols, int axis,
                                                        const int* __restrict__ idx,
                                                        double* __restrict__ dst) {
  const long long i = blockIdx.y;
  const double* srow = src + (axis == 0 ? (long long)idx[i] : i) * cols;
  double* drow = dst + i * cols;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < cols;
       j += (long long)gridDim.x * blockDim.x)
    drow[j] = srow[axis == 0 ? j : (long long)idx[j]];
}

extern "C" int pbx_gather_f64(pbx_ctx* ctx, const double* src, const int32_t* idx, int64_t n,
                              double* dst) {
  PBX_REQUIRE(ctx != nullptr, "pbx_gather_f64: null context");
  PBX_REQUIRE(n >= 0, "pbx_gather_f64: n must be >= 0");
  if (n == 0) return PBX_OK;
  PBX_REQUIRE(src && idx && dst, "pbx_gather_f64: null pointer");
  PBX_REQUIRE(src != dst, "pbx_gather_f64: src and dst must not alias");
  PBX_CUDA(cudaSetDevice(ctx->device));
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 16);
  gather_kernel<<<grid, 256, 0, ctx->stream>>>(src, idx, (long long)n, dst);
  PBX_LAUNCH_CHECK(ctx);
  return PBX_OK;
}

extern "C" int pbx_take_axis_f64(pbx_ctx* ctx, const double* src, int64_t rows, int64_t cols,
                                 int32_t axis, const int32_t* idx, double* dst) {
  PBX_REQUIRE(ctx != nullptr, "pbx_take_axis_f64: null context");
  PBX_REQUIRE(rows >= 0 && cols >= 0 && rows < 65536, "pbx_take_axis_f64: rows must be in [0, 65536)");
  PBX_REQUIRE(axis == 0 || axis == 1, "pbx_take_axis_f64: axis must be 0 or 1");
  if (rows == 0 || cols == 0) return PBX_OK;
  PBX_REQUIRE(src && idx && dst && src != dst, "pbx_take_axis_f64: null or aliased pointer");
  PBX_CUDA(cudaSetDevice(ctx->device));
  dim3 grid((unsigned)std::min<int64_t>((cols + 255) / 256, 64), (unsigned)rows);
  take_axis_kernel<<<grid, 256, 0, ctx->stream>>>(src, (long long)rows, (long long)cols, axis, idx,
                                                  dst);
  PBX_LAUNCH_CHECK(ctx);
  return PBX_OK;
}

// ===========================================================================
// Normalised cumulative probability: reduce-then-scan.
//   cp_tile_sum   tile totals              (read 8 B/element)
//   cp_scan_part  exclusive scan of the tile totals + grand total (one CTA)
//   cp_tile_scan  tile-local scan + offset, divided by max(tiny, total)
//                                          (read 8 B, write 8 B per element)
// Fixed association order everywhere -> identical bits run to run.
// ===========================================================================
#define CP_THREADS 256
#define CP_ITEMS 8
#define CP_TILE (CP_THREADS * CP_ITEMS)

__device__ __forceinline__ double cp_lin(double p, int log_pscale) {
  return log_pscale ? pbx_exp_logp(p) : p;
}

// block-wide exclusive prefix of one value per thread (thread order), also the block total
__device__ __forceinline__ double cp_block_excl(double v, double* s_w, double& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double inc = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, inc, off);
    if (lane >= off) inc += t;
  }
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  double wbase = 0.0, tot = 0.0;
#pragma unroll
  for (int w = 0; w < CP_THREADS / 32; ++w) {
    if (w == warp) wbase = tot;
    tot += s_w[w];
  }
  total = tot;
  __syncthreads();
  return wbase + (inc - v);
}

// block-wide exclusive running maximum of one value per thread (thread order); exact
__device__ __forceinline__ double cp_block_excl_max(double v, double* s_w, double& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double inc = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, inc, off);
    if (lane >= off) inc = fmax(inc, t);
  }
  if (lane == 31) s_w[warp] = inc;
  double ex = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane == 0) ex = 0.0;                      // all values are >= 0
  __syncthreads();
  double wbase = 0.0, tot = 0.0;
#pragma unroll
  for (int w = 0; w < CP_THREADS / 32; ++w) {
    if (w == warp) wbase = tot;
    tot = fmax(tot, s_w[w]);
  }
  total = tot;
  __syncthreads();
  return fmax(wbase, ex);
}

__device__ __forceinline__ void cp_load_tile(const double* __restrict__ prob, long long n,
                                             int log_pscale, double (&v)[CP_ITEMS]) {
  const long long i0 = (long long)blockIdx.x * CP_TILE + (long long)threadIdx.x * CP_ITEMS;
  if (i0 + CP_ITEMS <= n) {
    const double2* p2 = reinterpret_cast<const double2*>(prob + i0);
#pragma unroll
    for (int k = 0; k < CP_ITEMS / 2; ++k) {
      const double2 t = p2[k];
      v[2 * k] = cp_lin(t.x, log_pscale);
      v[2 * k + 1] = cp_lin(t.y, log_pscale);
    }
  } else {
#pragma unroll
    for (int k = 0; k < CP_ITEMS; ++k) v[k] = (i0 + k < n) ? cp_lin(prob[i0 + k], log_pscale) : 0.0;
  }
}

__global__ void __launch_bounds__(CP_THREADS) cp_tile_sum_kernel(const double* __restrict__ prob,
                                                                 long long n, int log_pscale,
                                                                 double* __restrict__ part) {
  __shared__ double s_w[CP_THREADS / 32];
  double v[CP_ITEMS];
  cp_load_tile(prob, n, log_pscale, v);
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < CP_ITEMS; ++k) s += v[k];
  double total;
  cp_block_excl(s, s_w, total);
  if (threadIdx.x == 0) part[blockIdx.x] = total;
}

// one CTA: part[t] <- sum_{u<t} part[u]; part[ntiles] = *total <- sum of all.
// A parallel sum scan is non-decreasing only to within a few ulp (each prefix has its
// own association); an exact running-maximum pass on top makes the offsets -- and with
// the clamps in cp_tile_scan the whole cumulative array -- truly non-decreasing, which
// np.digitize / a binary search rely on (pd.py:430).
__global__ void __launch_bounds__(CP_THREADS) cp_scan_part_kernel(double* __restrict__ part,
                                                                  int ntiles,
                                                                  double* __restrict__ total_ws,
                                                                  double* __restrict__ total_out) {
  __shared__ double s_w[CP_THREADS / 32];
  double carry = 0.0, carry_max = 0.0;
  for (int t0 = 0; t0 < ntiles; t0 += CP_TILE) {
    double v[CP_ITEMS], s = 0.0;
    const int i0 = t0 + threadIdx.x * CP_ITEMS;
#pragma unroll
    for (int k = 0; k < CP_ITEMS; ++k) {
      v[k] = (i0 + k < ntiles) ? part[i0 + k] : 0.0;
      s += v[k];
    }
    double total, tmax;
    double run = carry + cp_block_excl(s, s_w, total);
    double o[CP_ITEMS];
#pragma unroll
    for (int k = 0; k < CP_ITEMS; ++k) {
      o[k] = run;
      run += v[k];
    }
    const double floor_ = fmax(carry_max, cp_block_excl_max(o[CP_ITEMS - 1], s_w, tmax));
#pragma unroll
    for (int k = 0; k < CP_ITEMS; ++k)
      if (i0 + k < ntiles) part[i0 + k] = fmax(o[k], floor_);
    carry += total;
    carry_max = fmax(carry_max, tmax);
  }
  if (threadIdx.x == 0) {
    const double tot = fmax(carry, carry_max);
    part[ntiles] = tot;
    *total_ws = tot;
    if (total_out) *total_out = tot;
  }
}

__global__ void __launch_bounds__(CP_THREADS) cp_tile_scan_kernel(const double* __restrict__ prob,
                                                                  long long n, int log_pscale,
                                                                  const double* __restrict__ part,
                                                                  const double* __restrict__ total_ws,
                                                                  double* __restrict__ cum) {
  __shared__ double s_w[CP_THREADS / 32];
  double v[CP_ITEMS];
  cp_load_tile(prob, n, log_pscale, v);
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < CP_ITEMS; ++k) s += v[k];
  double total, tmax;
  const double lo = part[blockIdx.x], hi = part[blockIdx.x + 1];
  double run = lo + cp_block_excl(s, s_w, total);
  const double den = fmax(PBX_TINY, *total_ws);            // div_prob: pscales.py:219-236
  const long long i0 = (long long)blockIdx.x * CP_TILE + (long long)threadIdx.x * CP_ITEMS;
#pragma unroll
  for (int k = 0; k < CP_ITEMS; ++k) {
    run += v[k];
    v[k] = run;
  }
  // non-decreasing by construction: sequential inside a thread, running maximum across
  // the threads, clamped into [offset of this tile, offset of the next]
  const double floor_ = cp_block_excl_max(v[CP_ITEMS - 1], s_w, tmax);
#pragma unroll
  for (int k = 0; k < CP_ITEMS; ++k) {
    double c = fmin(fmax(v[k], floor_), hi);
    if (i0 + k == n - 1) c = *total_ws;                    // cum[-1] / cum[-1] = 1 exactly
    v[k] = c / den;
  }
  if (i0 + CP_ITEMS <= n) {
    double2* c2 = reinterpret_cast<double2*>(cum + i0);
#pragma unroll
    for (int k = 0; k < CP_ITEMS / 2; ++k) c2[k] = make_double2(v[2 * k], v[2 * k + 1]);
  } else {
#pragma unroll
    for (int k = 0; k < CP_ITEMS; ++k)
      if (i0 + k < n) cum[i0 + k] = v[k];
  }
}

static size_t scan_ws_bytes(int64_t n) {
  const size_t ntiles = (size_t)((n + CP_TILE - 1) / CP_TILE) + 2;
  // tile partials + grand total, or the expectation's per-CTA partials (<= 4096 x 9)
  return align256(ntiles * 8) + 256 + align256((size_t)4096 * 9 * 8);
}

extern "C" int pbx_scan_workspace_bytes(int64_t n, size_t* bytes) {
  PBX_REQUIRE(bytes != nullptr && n >= 0, "pbx_scan_workspace_bytes: bad argument");
  *bytes = scan_ws_bytes(n);
  return PBX_OK;
}

extern "C" int pbx_cumprob_f64(pbx_ctx* ctx, const double* prob, int64_t n, int32_t log_pscale,
                               double* cum, double* total, void* workspace,
                               size_t workspace_bytes) {
  PBX_REQUIRE(ctx != nullptr, "pbx_cumprob_f64: null context");
  PBX_REQUIRE(n >= 0, "pbx_cumprob_f64: n must be >= 0");
  if (n == 0) return PBX_OK;
  PBX_REQUIRE(prob && cum, "pbx_cumprob_f64: prob and cum are mandatory");
  PBX_REQUIRE(((uintptr_t)prob & 15) == 0 && ((uintptr_t)cum & 15) == 0,
              "pbx_cumprob_f64: prob and cum must be 16-byte aligned");
  PBX_REQUIRE(workspace != nullptr && workspace_bytes >= scan_ws_bytes(n),
              "pbx_cumprob_f64: workspace of %zu bytes needed (got %zu)", scan_ws_bytes(n),
              workspace_bytes);
  PBX_CUDA(cudaSetDevice(ctx->device));
  const int64_t ntiles64 = (n + CP_TILE - 1) / CP_TILE;
  PBX_REQUIRE(ntiles64 < ((int64_t)1 << 31), "pbx_cumprob_f64: n too large");
  const int ntiles = (int)ntiles64;
  double* part = static_cast<double*>(workspace);
  double* total_ws = reinterpret_cast<double*>(static_cast<unsigned char*>(workspace) +
                                               align256((size_t)(ntiles + 2) * 8));
  cp_tile_sum_kernel<<<ntiles, CP_THREADS, 0, ctx->stream>>>(prob, (long long)n, log_pscale, part);
  PBX_LAUNCH_CHECK(ctx);
  cp_scan_part_kernel<<<1, CP_THREADS, 0, ctx->stream>>>(part, ntiles, total_ws, total);
  PBX_LAUNCH_CHECK(ctx);
  cp_tile_scan_kernel<<<ntiles, CP_THREADS, 0, ctx->stream>>>(prob, (long long)n, log_pscale, part,
                                                             total_ws, cum);
  PBX_LAUNCH_CHECK(ctx);
  return PBX_OK;
}

// np.digitize(q, cum) for increasing bins = number of bins <= q (upper bound)
struct DigQ { double q[64]; };
__global__ void digitize_kernel(const double* __restrict__ cum, long long n, const DigQ qs, int nq,
                                long long* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nq) return;
  const double q = qs.q[k];
  long long lo = 0, hi = n;                      // first index with cum[i] > q
  while (lo < hi) {
    const long long mid = lo + ((hi - lo) >> 1);
    if (cum[mid] <= q) lo = mid + 1; else hi = mid;
  }
  out[k] = (lo - 1 > 0) ? lo - 1 : 0;
}

extern "C" int pbx_digitize_f64(pbx_ctx* ctx, const double* cum, int64_t n, const double* q,
                                int32_t nq, int64_t* idx_out) {
  PBX_REQUIRE(ctx != nullptr, "pbx_digitize_f64: null context");
  PBX_REQUIRE(n >= 1 && cum && q && idx_out, "pbx_digitize_f64: bad argument");
  PBX_REQUIRE(nq >= 1 && nq <= 64, "pbx_digitize_f64: nq must be in 1..64 (got %d)", nq);
  PBX_CUDA(cudaSetDevice(ctx->device));
  DigQ qs;
  for (int k = 0; k < 64; ++k) qs.q[k] = k < nq ? q[k] : 0.0;
  digitize_kernel<<<1, 64, 0, ctx->stream>>>(cum, (long long)n, qs, nq,
                                             reinterpret_cast<long long*>(idx_out));
  PBX_LAUNCH_CHECK(ctx);
  return PBX_OK;
}

// ===========================================================================
// Expectation sums: one pass over prob, per-CTA partials in a fixed partition,
// then one CTA adds the partials in index order.
// ===========================================================================
#define EX_THREADS 256
#define EX_MAXV 4
#define EX_NACC (1 + 2 * EX_MAXV)

template <int KR, int KC>
__global__ void __launch_bounds__(EX_THREADS) ex_partial_kernel(
    const double* __restrict__ prob, long long rows, long long cols, int log_pscale,
    const double* __restrict__ row_vals, const double* __restrict__ col_vals,
    long long per_cta, double* __restrict__ part) {
  __shared__ double s_red[EX_THREADS / 32][1 + KR + KC];
  const long long n = rows * cols;
  const long long e0 = (long long)blockIdx.x * per_cta;
  const long long e1 = min(n, e0 + per_cta);
  double acc[1 + KR + KC];
#pragma unroll
  for (int a = 0; a < 1 + KR + KC; ++a) acc[a] = 0.0;
  const bool small = n < 0xFFFFFFFFll;
  auto term = [&](long long e, double praw) {
    const double p = cp_lin(praw, log_pscale);
    long long i = 0, j = e;
    if (rows > 1) {
      if (small) {
        const unsigned ii = (unsigned)e / (unsigned)cols;
        i = ii;
        j = (unsigned)e - ii * (unsigned)cols;
      } else {
        i = e / cols;
        j = e - i * cols;
      }
    }
    acc[0] += p;
#pragma unroll
    for (int k = 0; k < KR; ++k) acc[1 + k] = fma(p, row_vals[k * rows + i], acc[1 + k]);
#pragma unroll
    for (int k = 0; k < KC; ++k) acc[1 + KR + k] = fma(p, col_vals[k * cols + j], acc[1 + KR + k]);
  };
  // four independent cells in flight per thread (the loads are issued before the exps)
  long long e = e0 + threadIdx.x;
  for (; e + 3 * EX_THREADS < e1; e += 4 * EX_THREADS) {
    double pr[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) pr[u] = prob[e + u * EX_THREADS];
#pragma unroll
    for (int u = 0; u < 4; ++u) term(e + u * EX_THREADS, pr[u]);
  }
  for (; e < e1; e += EX_THREADS) term(e, prob[e]);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < 1 + KR + KC; ++a) {
    double v = acc[a];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (lane == 0) s_red[warp][a] = v;
  }
  __syncthreads();
  if (threadIdx.x < 1 + KR + KC) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < EX_THREADS / 32; ++w) v += s_red[w][threadIdx.x];
    part[(size_t)blockIdx.x * EX_NACC + threadIdx.x] = v;
  }
}

__global__ void __launch_bounds__(256) ex_final_kernel(const double* __restrict__ part, int nparts,
                                                       int nacc, double* __restrict__ out) {
  // thread a (< nacc) of warp w adds a strided share in index order; then a fixed tree
  __shared__ double s[8][EX_NACC];
  const int a = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (a < nacc) {
    double v = 0.0;
    for (int p = w; p < nparts; p += 8) v += part[(size_t)p * EX_NACC + a];
    s[w][a] = v;
  }
  __syncthreads();
  if (w == 0 && a < nacc) {
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += s[k][a];
    out[a] = v;
  }
}

template <int KR>
static void ex_launch_kc(int kc, int grid, cudaStream_t st, const double* prob, long long rows,
                         long long cols, int lg, const double* rv, const double* cv,
                         long long per_cta, double* part) {
  switch (kc) {
    case 0: ex_partial_kernel<KR, 0><<<grid, EX_THREADS, 0, st>>>(prob, rows, cols, lg, rv, cv, per_cta, part); break;
    case 1: ex_partial_kernel<KR, 1><<<grid, EX_THREADS, 0, st>>>(prob, rows, cols, lg, rv, cv, per_cta, part); break;
    case 2: ex_partial_kernel<KR, 2><<<grid, EX_THREADS, 0, st>>>(prob, rows, cols, lg, rv, cv, per_cta, part); break;
    case 3: ex_partial_kernel<KR, 3><<<grid, EX_THREADS, 0, st>>>(prob, rows, cols, lg, rv, cv, per_cta, part); break;
    default: ex_partial_kernel<KR, 4><<<grid, EX_THREADS, 0, st>>>(prob, rows, cols, lg, rv, cv, per_cta, part); break;
  }
}

extern "C" int pbx_expectation_f64(pbx_ctx* ctx, const double* prob, int64_t rows, int64_t cols,
                                   int32_t log_pscale, const double* row_vals, int32_t n_row_vals,
                                   const double* col_vals, int32_t n_col_vals, double* out,
                                   void* workspace, size_t workspace_bytes) {
  PBX_REQUIRE(ctx != nullptr, "pbx_expectation_f64: null context");
  PBX_REQUIRE(rows >= 1 && cols >= 1 && prob && out, "pbx_expectation_f64: bad argument");
  PBX_REQUIRE(n_row_vals >= 0 && n_row_vals <= EX_MAXV && n_col_vals >= 0 && n_col_vals <= EX_MAXV,
              "pbx_expectation_f64: at most %d value arrays per axis", EX_MAXV);
  PBX_REQUIRE((n_row_vals == 0 || row_vals) && (n_col_vals == 0 || col_vals),
              "pbx_expectation_f64: value arrays missing");
  const int64_t n = rows * cols;
  PBX_REQUIRE(workspace != nullptr && workspace_bytes >= scan_ws_bytes(n),
              "pbx_expectation_f64: workspace of %zu bytes needed (got %zu)", scan_ws_bytes(n),
              workspace_bytes);
  PBX_CUDA(cudaSetDevice(ctx->device));
  // fixed partition: up to 8 CTAs per SM, each a whole number of 2048-element blocks
  int64_t grid = std::min<int64_t>((n + 2047) / 2048, std::min<int64_t>(4096, (int64_t)ctx->sm_count * 8));
  int64_t per_cta = ((n + grid - 1) / grid + 2047) / 2048 * 2048;
  grid = (n + per_cta - 1) / per_cta;
  double* part = static_cast<double*>(workspace);
  const int lg = log_pscale;
  switch (n_row_vals) {
    case 0: ex_launch_kc<0>(n_col_vals, (int)grid, ctx->stream, prob, rows, cols, lg, row_vals, col_vals, per_cta, part); break;
    case 1: ex_launch_kc<1>(n_col_vals, (int)grid, ctx->stream, prob, rows, cols, lg, row_vals, col_vals, per_cta, part); break;
    case 2: ex_launch_kc<2>(n_col_vals, (int)grid, ctx->stream, prob, rows, cols, lg, row_vals, col_vals, per_cta, part); break;
    case 3: ex_launch_kc<3>(n_col_vals, (int)grid, ctx->stream, prob, rows, cols, lg, row_vals, col_vals, per_cta, part); break;
    default: ex_launch_kc<4>(n_col_vals, (int)grid, ctx->stream, prob, rows, cols, lg, row_vals, col_vals, per_cta, part); break;
  }
  PBX_LAUNCH_CHECK(ctx);
  ex_final_kernel<<<1, 256, 0, ctx->stream>>>(part, (int)grid, 1 + n_row_vals + n_col_vals, out);
  PBX_LAUNCH_CHECK(ctx);
  return PBX_OK;
}

// ===========================================================================
// Binary PD algebra with broadcasting: the product rule (PD.__mul__ -> pd_utils.product
// -> pscales.prod_rule, probayes/pscales.py:160-216) and the safe division
// (PD.__truediv__ -> pscales.div_prob, pscales.py:219-236) of two probability arrays
// whose shapes are [rows or 1][cols or 1].  One streaming pass, 8 B out per cell.
// ===========================================================================
__global__ void __launch_bounds__(256) pd_binary_kernel(int op, const double* __restrict__ a,
                                                        long long a_rs, long long a_cs, int a_log,
                                                        const double* __restrict__ b,
                                                        long long b_rs, long long b_cs, int b_log,
                                                        long long rows, long long cols, int out_log,
                                                        double* __restrict__ out) {
  const long long n = rows * cols;
  const bool small = n < 0xFFFFFFFFll;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (long long)gridDim.x * blockDim.x) {
    long long i, j;
    if (small) {
      const unsigned ii = (unsigned)e / (unsigned)cols;
      i = ii;
      j = (unsigned)e - ii * (unsigned)cols;
    } else {
      i = e / cols;
      j = e - i * cols;
    }
    const double av = a[i * a_rs + j * a_cs], bv = b[i * b_rs + j * b_cs];
    double r;
    if (op == 0) {
      // product rule: log space as soon as one factor is in log pscale (clamped log of
      // the linear one), plain product otherwise
      if (out_log) r = (a_log ? av : pbx_log_prob(av)) + (b_log ? bv : pbx_log_prob(bv));
      else r = av * bv;
    } else {
      // div_prob: both to linear, num / max(tiny, den), back to the output pscale
      const double num = a_log ? pbx_exp_logp(av) : av;
      const double den = b_log ? pbx_exp_logp(bv) : bv;
      const double q = num / fmax(PBX_TINY, den);
      r = out_log ? pbx_log_prob(q) : q;
    }
    out[e] = r;
  }
}

extern "C" int pbx_pd_binary_f64(pbx_ctx* ctx, int32_t op, const double* a, int64_t a_rows,
                                 int64_t a_cols, int32_t a_log, const double* b, int64_t b_rows,
                                 int64_t b_cols, int32_t b_log, int64_t rows, int64_t cols,
                                 int32_t out_log, double* out) {
  PBX_REQUIRE(ctx != nullptr, "pbx_pd_binary_f64: null context");
  PBX_REQUIRE(op == 0 || op == 1, "pbx_pd_binary_f64: op must be 0 (product) or 1 (division)");
  PBX_REQUIRE(rows >= 0 && cols >= 0, "pbx_pd_binary_f64: negative shape");
  if (rows == 0 || cols == 0) return PBX_OK;
  PBX_REQUIRE(a && b && out, "pbx_pd_binary_f64: null pointer");
  PBX_REQUIRE((a_rows == 1 || a_rows == rows) && (a_cols == 1 || a_cols == cols) &&
                  (b_rows == 1 || b_rows == rows) && (b_cols == 1 || b_cols == cols),
              "pbx_pd_binary_f64: operand shapes must be [rows or 1][cols or 1]");
  PBX_REQUIRE(op == 1 || (out_log != 0) == (a_log != 0 || b_log != 0),
              "pbx_pd_binary_f64: the product is in log pscale iff a factor is "
              "(pscales.py:170-172)");
  PBX_CUDA(cudaSetDevice(ctx->device));
  const long long a_rs = a_rows == 1 ? 0 : a_cols, a_cs = a_cols == 1 ? 0 : 1;
  const long long b_rs = b_rows == 1 ? 0 : b_cols, b_cs = b_cols == 1 ? 0 : 1;
  const int64_t n = rows * cols;
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 16);
  pd_binary_kernel<<<grid, 256, 0, ctx->stream>>>(op, a, a_rs, a_cs, a_log, b, b_rs, b_cs, b_log,
                                                  (long long)rows, (long long)cols, out_log, out);
  PBX_LAUNCH_CHECK(ctx);
  return PBX_OK;
}

// ===========================================================================
// Box sampler of ordinary Monte Carlo random sampling
// ===========================================================================
struct BoxConst {
  double ulo[8], ulen[8];
  int lg[8];
};

__global__ void __launch_bounds__(256) box_sample_kernel(int P, long long T, const BoxConst bc,
                                                         unsigned long long seed, long long t0,
                                                         const double* __restrict__ inj,
                                                         double* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  for (int s = 0; s < (P + 1) / 2; ++s) {
    double r0, r1 = 0.0;
    if (inj) {
      r0 = inj[t * P + 2 * s];
      if (2 * s + 1 < P) r1 = inj[t * P + 2 * s + 1];
    } else {
      const pbx_u4 w = pbx_block(seed, (uint64_t)(t0 + t), 0u, (uint32_t)s);
      r0 = pbx_u52(w.x, w.y);
      r1 = pbx_u32(w.z);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = 2 * s + h;
      if (j >= P) break;
      // np.random.uniform(lo, hi) = lo + (hi - lo) * r  in ufun space, then ufun^-1
      // (separate multiply and add, as numpy: no FMA contraction)
      const double u = __dadd_rn(bc.ulo[j], __dmul_rn(bc.ulen[j], h ? r1 : r0));
      out[(long long)j * T + t] = bc.lg[j] ? exp(u) : u;
    }
  }
}

extern "C" int pbx_box_sample(pbx_ctx* ctx, int32_t n_params, int64_t n_samples, const double* lims,
                              const int32_t* log_ufun, uint64_t seed, int64_t sample0,
                              const double* inj_unif, double* out) {
  PBX_REQUIRE(ctx != nullptr, "pbx_box_sample: null context");
  PBX_REQUIRE(n_params >= 1 && n_params <= 8, "pbx_box_sample: n_params must be in 1..8 (got %d)",
              n_params);
  PBX_REQUIRE(n_samples >= 0 && sample0 >= 0, "pbx_box_sample: n_samples/sample0 must be >= 0");
  if (n_samples == 0) return PBX_OK;
  PBX_REQUIRE(lims && out, "pbx_box_sample: lims and out are mandatory");
  BoxConst bc;
  for (int j = 0; j < 8; ++j) {
    bc.ulo[j] = bc.ulen[j] = 0.0;
    bc.lg[j] = 0;
    if (j >= n_params) continue;
    const double lo = lims[2 * j], hi = lims[2 * j + 1];
    bc.lg[j] = log_ufun ? (log_ufun[j] != 0) : 0;
    PBX_REQUIRE(std::isfinite(lo) && std::isfinite(hi) && hi >= lo,
                "pbx_box_sample: parameter %d needs finite limits (variable.py:572-574)", j);
    PBX_REQUIRE(!bc.lg[j] || lo > 0.0, "pbx_box_sample: log ufun needs positive limits (param %d)", j);
    const double ulo = bc.lg[j] ? log(lo) : lo, uhi = bc.lg[j] ? log(hi) : hi;
    bc.ulo[j] = ulo;
    bc.ulen[j] = uhi - ulo;
  }
  PBX_CUDA(cudaSetDevice(ctx->device));
  const int grid = (int)((n_samples + 255) / 256);
  box_sample_kernel<<<grid, 256, 0, ctx->stream>>>(n_params, (long long)n_samples, bc, seed,
                                                   (long long)sample0, inj_unif, out);
  PBX_LAUNCH_CHECK(ctx);
  return PBX_OK;
}


// ===========================================================================
// Ordinary Monte Carlo with REJECTION sampling (examples/omc/omc_rejection_sp_circle.py:
// 26-39; SP.next with a proposal density, probayes/sp.py:221-258, sd.py:228-250,
// rf.py:584-602).  Per sample: every variable afresh from its box (uniform in ufun space,
// variable.py:558-583), the proposal density q there, the target p, the score s (p or the
// safe ratio p / q), one threshold uniform t in [t_lo, t_hi) and the update flag s >= t.
// RNG: the box draws are those of pbx_box_sample (Philox block (seed, sample, 0, slot));
// the threshold is u52 of block (seed, sample, 1, 0).  Injected: [T][P + 1] per sample
// (draws in variable order, then the threshold), the reference's call order.
// ===========================================================================
struct RejConst {
  double ulo[8], ulen[8];
  int lg[8];
  double centre[8], loc[8], scale[8], radius, t_lo, t_hi;
  int target_kind, prop_kind, score_mode;
};

__global__ void __launch_bounds__(256) rejection_kernel(int P, long long T, const RejConst rc,
                                                        unsigned long long seed, long long t0,
                                                        const double* __restrict__ inj,
                                                        double* __restrict__ theta,
                                                        double* __restrict__ op,
                                                        double* __restrict__ oq,
                                                        double* __restrict__ os,
                                                        double* __restrict__ ot,
                                                        unsigned char* __restrict__ ou) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  double x[8];
  for (int s = 0; s < (P + 1) / 2; ++s) {
    double r0, r1 = 0.0;
    if (inj) {
      r0 = inj[t * (P + 1) + 2 * s];
      if (2 * s + 1 < P) r1 = inj[t * (P + 1) + 2 * s + 1];
    } else {
      const pbx_u4 w = pbx_block(seed, (uint64_t)(t0 + t), 0u, (uint32_t)s);
      r0 = pbx_u52(w.x, w.y);
      r1 = pbx_u32(w.z);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = 2 * s + h;
      if (j >= P) break;
      const double u = __dadd_rn(rc.ulo[j], __dmul_rn(rc.ulen[j], h ? r1 : r0));
      x[j] = rc.lg[j] ? exp(u) : u;
      theta[(long long)j * T + t] = x[j];
    }
  }
  // proposal density: product of normal pdfs as scipy evaluates them (exp(-z^2/2) /
  // sqrt(2 pi) / scale), or of the box-uniform densities 1 / length
  double q = 1.0;
  for (int j = 0; j < P; ++j) {
    if (rc.prop_kind == 0) {
      const double z = (x[j] - rc.loc[j]) / rc.scale[j];
      q *= (exp(-(z * z) / 2.0) / 2.50662827463100050242) / rc.scale[j];
    } else {
      q *= 1.0 / rc.ulen[j];
    }
  }
  // target: indicator of a ball (1.0 / 0.0)
  double ss = 0.0;
  for (int j = 0; j < P; ++j) {
    const double dlt = x[j] - rc.centre[j];
    ss += __dmul_rn(dlt, dlt);
  }
  const double p = (ss <= rc.radius * rc.radius) ? 1.0 : 0.0;
  const double sc = rc.score_mode == 0 ? p : p / fmax(PBX_TINY, q);    // div_prob
  double r;
  if (inj) {
    r = inj[t * (P + 1) + P];
  } else {
    const pbx_u4 w = pbx_block(seed, (uint64_t)(t0 + t), 1u, 0u);
    r = pbx_u52(w.x, w.y);
  }
  const double th = __dadd_rn(rc.t_lo, __dmul_rn(rc.t_hi - rc.t_lo, r));
  op[t] = p;
  oq[t] = q;
  os[t] = sc;
  ot[t] = th;
  ou[t] = (sc >= th) ? 1 : 0;
}

extern "C" int pbx_rejection_sample(pbx_ctx* ctx, const pbx_rejection_params* p) {
  PBX_REQUIRE(ctx != nullptr && p != nullptr, "pbx_rejection_sample: null argument");
  PBX_REQUIRE(p->n_params >= 1 && p->n_params <= 8,
              "pbx_rejection_sample: n_params must be in 1..8 (got %d)", p->n_params);
  PBX_REQUIRE(p->n_samples >= 0 && p->sample0 >= 0, "pbx_rejection_sample: negative size");
  PBX_REQUIRE(p->target_kind == 0, "pbx_rejection_sample: unknown target_kind %d", p->target_kind);
  PBX_REQUIRE(p->prop_kind == 0 || p->prop_kind == 1,
              "pbx_rejection_sample: unknown prop_kind %d", p->prop_kind);
  PBX_REQUIRE(p->score_mode == 0 || p->score_mode == 1,
              "pbx_rejection_sample: unknown score_mode %d", p->score_mode);
  if (p->n_samples == 0) return PBX_OK;
  PBX_REQUIRE(p->out_theta && p->out_p && p->out_q && p->out_s && p->out_t && p->out_u,
              "pbx_rejection_sample: every output buffer is mandatory");
  RejConst rc;
  for (int j = 0; j < 8; ++j) {
    rc.ulo[j] = rc.ulen[j] = rc.centre[j] = rc.loc[j] = 0.0;
    rc.scale[j] = 1.0;
    rc.lg[j] = 0;
    if (j >= p->n_params) continue;
    const double lo = p->lims[j][0], hi = p->lims[j][1];
    rc.lg[j] = p->log_ufun[j] != 0;
    PBX_REQUIRE(std::isfinite(lo) && std::isfinite(hi) && hi >= lo,
                "pbx_rejection_sample: variable %d needs finite limits (variable.py:572-574)", j);
    PBX_REQUIRE(!rc.lg[j] || lo > 0.0, "pbx_rejection_sample: log ufun needs positive limits");
    const double ulo = rc.lg[j] ? log(lo) : lo, uhi = rc.lg[j] ? log(hi) : hi;
    rc.ulo[j] = ulo;
    rc.ulen[j] = uhi - ulo;
    rc.centre[j] = p->target_centre[j];
    rc.loc[j] = p->prop_loc[j];
    rc.scale[j] = p->prop_scale[j];
    PBX_REQUIRE(p->prop_kind != 0 || rc.scale[j] > 0.0,
                "pbx_rejection_sample: proposal scale %d must be positive", j);
  }
  rc.radius = p->target_radius;
  rc.t_lo = p->thresh_lo;
  rc.t_hi = p->thresh_hi;
  rc.target_kind = p->target_kind;
  rc.prop_kind = p->prop_kind;
  rc.score_mode = p->score_mode;
  PBX_CUDA(cudaSetDevice(ctx->device));
  const int grid = (int)((p->n_samples + 255) / 256);
  rejection_kernel<<<grid, 256, 0, ctx->stream>>>(p->n_params, (long long)p->n_samples, rc,
                                                  p->seed, (long long)p->sample0, p->inj_unif,
                                                  p->out_theta, p->out_p, p->out_q, p->out_s,
                                                  p->out_t, p->out_u);
  PBX_LAUNCH_CHECK(ctx);
  return PBX_OK;
}
