// pbx_ndtri.cuh -- table-driven inverse normal CDF for the Gibbs draws (K5).
//
// The reference draws a truncated conditional normal by inversion (vtypes.py:186 ->
// scipy.stats.norm.ppf == ndtri): x = ndtri(cdf_lo + (cdf_hi - cdf_lo) r) * stdv + mean.
// CUDA's normcdfinv is ~500 static SASS instructions with a central / tail split on which a
// warp always diverges (P(all 32 lanes central) < 1 %): with Philox it was 60 % of the Gibbs
// kernel's instructions.  Here ndtri(p), p = min(u, 1 - u), is a degree-5 polynomial in the
// mantissa offset of p on 64 segments per binade, 64 binades (p >= 2^-64; u52 uniforms are
// >= 2^-53): the segment index is bits 62..46 of p, the local variable v = mantissa - segment
// centre is exact (|v| <= 2^-7), six coefficients = one 48-byte row.  ndtri is analytic in p
// away from 0 with radius of convergence p, so the segment-relative width 1/64 gives the
// interpolant at Chebyshev nodes an error ~ (1/256)^6: measured <= 7e-16 relative to
// max(|x|, 1e-3) against 40-digit values over all binades (<= 3.4e-16 below 1/8).  (A
// degree-7 table on 32 segments measures 2.3e-16; its eight loads per draw instead of six
// cost more shared-memory bandwidth than the Gibbs kernel has.)  The 192 KB table lives in
// global memory and is read through L1.  Outside the table (p < 2^-64, u <= 0, u >= 1, NaN) a cold out-of-line call to
// normcdfinv keeps the reference's limits (ndtri(0) = -inf, ndtri(1) = +inf).
//
// The table is generated on the host at context initialisation in long double (x87 64-bit
// mantissa): Acklam's rational start + two Halley steps on erfcl, Chebyshev-node interpolation,
// monomial coefficients by Gaussian elimination with partial pivoting.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#define NDT_EMIN 959                 // smallest tabulated biased exponent: p >= 2^-64
#define NDT_BINADES 64               // biased exponents 959 .. 1022
#define NDT_SEG_BITS 6               // leading mantissa bits that select the segment
#define NDT_SEGS (1 << NDT_SEG_BITS)
#define NDT_NCOEF 6                  // degree 5
#define NDT_ROWS (NDT_BINADES * NDT_SEGS)
#define NDT_SHIFT (20 - NDT_SEG_BITS) // segment index = high word >> NDT_SHIFT

#ifdef __CUDACC__
#define NDT_HD __host__ __device__
#else
#define NDT_HD
#endif

// ---- evaluation: identical arithmetic on host and device --------------------------------
NDT_HD inline int ndt_hi(double x) {
#ifdef __CUDA_ARCH__
  return __double2hiint(x);
#else
  uint64_t b; memcpy(&b, &x, 8); return (int)(b >> 32);
#endif
}
NDT_HD inline double ndt_make(int hi, uint32_t lo) {
#ifdef __CUDA_ARCH__
  return __hiloint2double(hi, (int)lo);
#else
  uint64_t b = ((uint64_t)(uint32_t)hi << 32) | lo; double x; memcpy(&x, &b, 8); return x;
#endif
}
NDT_HD inline uint32_t ndt_lo(double x) {
#ifdef __CUDA_ARCH__
  return (uint32_t)__double2loint(x);
#else
  uint64_t b; memcpy(&b, &x, 8); return (uint32_t)b;
#endif
}
// segment of p (row of the table), or >= NDT_ROWS when p is outside the table
NDT_HD inline unsigned ndt_segment(double p) {
  return (unsigned)(ndt_hi(p) >> NDT_SHIFT) - ((unsigned)NDT_EMIN << NDT_SEG_BITS);
}
// polynomial of row c[0..5] at p (p inside the row's segment)
NDT_HD inline double ndt_poly(const double* c, double p) {
  const int hm = (ndt_hi(p) & 0x000FFFFF) | 0x3FF00000;            // mantissa m in [1, 2)
  const double m = ndt_make(hm, ndt_lo(p));
  const double cen = ndt_make((hm & ~((1 << NDT_SHIFT) - 1)) | (1 << (NDT_SHIFT - 1)), 0u);
  const double v = m - cen;                 // exact, |v| <= 2^-(NDT_SEG_BITS + 1); cen = centre
  double x = fma(c[5], v, c[4]);
  x = fma(x, v, c[3]);
  x = fma(x, v, c[2]);
  x = fma(x, v, c[1]);
  return fma(x, v, c[0]);
}

// ---- host: table generation --------------------------------------------------------------
static inline long double ndt_ndtri_l(long double p) {            // 0 < p < 1
  if (p > 0.5L) return -ndt_ndtri_l(1.0L - p);
  // Acklam's rational approximation (relative error 1.15e-9), lower half
  static const double a[6] = {-3.969683028665376e+01, 2.209460984245205e+02,
                              -2.759285104469687e+02, 1.383577518672690e+02,
                              -3.066479806614716e+01, 2.506628277459239e+00};
  static const double b[5] = {-5.447609879822406e+01, 1.615858368580409e+02,
                              -1.556989798598866e+02, 6.680131188771972e+01,
                              -1.328068155288572e+01};
  static const double c[6] = {-7.784894002430293e-03, -3.223964580411365e-01,
                              -2.400758277161838e+00, -2.549732539343734e+00,
                              4.374664141464968e+00, 2.938163982698783e+00};
  static const double d[4] = {7.784695709041462e-03, 3.224671290700398e-01,
                              2.445134137142996e+00, 3.754408661907416e+00};
  long double x;
  if (p < 0.02425L) {
    const long double q = sqrtl(-2.0L * logl(p));
    x = (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
        ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1.0L);
  } else {
    const long double q = p - 0.5L, r = q * q;
    x = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
        (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1.0L);
  }
  const long double sqrt2pi = 2.50662827463100050241576528481104525L;
  const long double rsqrt2 = 0.70710678118654752440084436210484904L;
  for (int it = 0; it < 3; ++it) {                                 // Halley on Phi(x) - p
    const long double e = 0.5L * erfcl(-x * rsqrt2) - p;
    const long double u = e * sqrt2pi * expl(0.5L * x * x);
    x -= u / (1.0L + 0.5L * x * u);
  }
  return x;
}

// fills tab[NDT_ROWS][NDT_NCOEF]
static inline void ndt_build_table(double* tab) {
  const int n = NDT_NCOEF;
  const long double pi = 3.14159265358979323846264338327950288L;
  long double t[NDT_NCOEF];
  for (int i = 0; i < n; ++i) t[i] = cosl(pi * (2 * i + 1) / (2.0L * n));
  const long double h = 1.0L / (2 * NDT_SEGS);                      // half a segment
  for (int be = 0; be < NDT_BINADES; ++be) {
    const long double scale = ldexpl(1.0L, NDT_EMIN + be - 1023);
    for (int k = 0; k < NDT_SEGS; ++k) {
      const long double cen = 1.0L + (k + 0.5L) / NDT_SEGS;
      long double A[NDT_NCOEF][NDT_NCOEF + 1];
      for (int i = 0; i < n; ++i) {
        long double pw = 1.0L;
        for (int j = 0; j < n; ++j) { A[i][j] = pw; pw *= t[i]; }
        A[i][n] = ndt_ndtri_l((cen + h * t[i]) * scale);
      }
      for (int col = 0; col < n; ++col) {                          // partial pivoting
        int piv = col;
        for (int r = col + 1; r < n; ++r)
          if (fabsl(A[r][col]) > fabsl(A[piv][col])) piv = r;
        if (piv != col)
          for (int j = 0; j <= n; ++j) { long double s = A[col][j]; A[col][j] = A[piv][j]; A[piv][j] = s; }
        for (int r = col + 1; r < n; ++r) {
          const long double f = A[r][col] / A[col][col];
          for (int j = col; j <= n; ++j) A[r][j] -= f * A[col][j];
        }
      }
      long double sol[NDT_NCOEF];
      for (int i = n - 1; i >= 0; --i) {
        long double s = A[i][n];
        for (int j = i + 1; j < n; ++j) s -= A[i][j] * sol[j];
        sol[i] = s / A[i][i];
      }
      long double hp = 1.0L;                                       // coefficients in v = h t
      double* row = tab + ((size_t)be * NDT_SEGS + k) * NDT_NCOEF;
      for (int j = 0; j < n; ++j) { row[j] = (double)(sol[j] / hp); hp *= h; }
    }
  }
}

// host mirror of the device evaluation (tests without a GPU); outside the table: NaN
static inline double ndt_eval_host(const double* tab, double u) {
  const bool upper = u > 0.5;
  const double p = upper ? 1.0 - u : u;
  const unsigned seg = ndt_segment(p);
  if (seg >= (unsigned)NDT_ROWS) return NAN;
  const double x = ndt_poly(tab + (size_t)seg * NDT_NCOEF, p);
  return upper ? -x : x;
}
