// pirb_device.cuh — 64-bit modular arithmetic and the shared-memory NTT core for sm_100a.
//
// Arithmetic contract (what the reference gets from SEAL 3.5.6):
// every public result is the canonical representative in [0,q).  Internally we use
// Harvey lazy butterflies ([0,4q) forward, [0,2q) inverse), Shoup multiplication for
// fixed twiddles and 128-bit lazy accumulation + one Barrett reduction for MACs.
// All of it is exact integer arithmetic, so results are bit-identical to any other
// correct implementation (the CPU oracle in particular).
#pragma once
#include <cuda_runtime.h>

#include "pirb_common.h"

namespace pirb {

// ------------------------------------------------------------------------------------------
// scalar helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 csub(u64 x, u64 q) { return x >= q ? x - q : x; }
__device__ __forceinline__ u64 addmod(u64 a, u64 b, u64 q) { return csub(a + b, q); }
__device__ __forceinline__ u64 submod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }
__device__ __forceinline__ u64 negmod(u64 a, u64 q) { return a ? q - a : 0; }

// x*w mod q in [0,2q) given ws = floor(w*2^64/q); valid for any 64-bit x, q < 2^63
__device__ __forceinline__ u64 shoup_lazy(u64 x, u64 w, u64 ws, u64 q) {
  u64 h = __umul64hi(x, ws);
  return x * w - h * q;
}
__device__ __forceinline__ u64 shoup(u64 x, u64 w, u64 ws, u64 q) { return csub(shoup_lazy(x, w, ws, q), q); }

// (hi:lo) += a*b
__device__ __forceinline__ void mac128(u64& lo, u64& hi, u64 a, u64 b) {
  u64 pl = a * b, ph = __umul64hi(a, b);
  asm("add.cc.u64 %0, %0, %2;\n\taddc.u64 %1, %1, %3;" : "+l"(lo), "+l"(hi) : "l"(pl), "l"(ph));
}

__device__ __forceinline__ u64 barrett128(u64 lo, u64 hi, u64 q, u64 r_hi, u64 r_lo);
__device__ __forceinline__ double f64_modmul(double y, double w, double wi, double q);
__device__ __forceinline__ double f64_canon(double v, double q, double qinv);
__device__ __forceinline__ u64 f64_to_u64_exact(double d);

// ------------------------------------------------------------------------------------------
// Lazy multiply-accumulate chains  sum_i a_i*b_i  (mod q taken once at the end).
//   MAC_WIDE : any 64-bit operands, one 128-bit accumulator (64x64->128 products with carries).
//   MAC_INT24: operands below 2^48 split into 24-bit halves, Karatsuba: three carry-free 64-bit partial sums,
//              3 IMAD.WIDE.U32 per product, <= 2^14 terms per chain.
//   MAC_FP64 : operands below 2^44 (all BFVDefault primes up to N=8192) split into h<=22-bit halves held as
//              doubles; 3 DFMA per product on the FP64 pipe, <= 2^(51-2h) terms per chain.
// All three are exact, so the reduced result is identical.
// ------------------------------------------------------------------------------------------
constexpr u32 PIRB_SMALL_MAX_TERMS = 1u << 14;
enum { MAC_WIDE = 0, MAC_INT24 = 1, MAC_FP64 = 2 };

template <int MODE>
struct Opnd;
template <int MODE>
struct Acc;

// ---- MAC_WIDE: any 64-bit operands, 128-bit accumulator ----
template <>
struct Opnd<MAC_WIDE> {
  u64 v;
  __device__ __forceinline__ Opnd(u64 x, int) : v(x) {}
};
template <>
struct Acc<MAC_WIDE> {
  u64 lo = 0, hi = 0;
  __device__ __forceinline__ void mac(const Opnd<MAC_WIDE>& a, const Opnd<MAC_WIDE>& b) { mac128(lo, hi, a.v, b.v); }
  __device__ __forceinline__ u64 reduce(const ModC& m, int) const {
    return barrett128(lo, hi, m.q, m.ratio_hi, m.ratio_lo);
  }
};

// value = s0 + s1*2^h + s2*2^(2h) (s1 = sk - s0 - s2: Karatsuba cross terms) reduced mod q
__device__ __forceinline__ u64 karatsuba_reduce(u64 s0, u64 sk, u64 s2, int h, const ModC& m) {
  const u64 s1 = sk - s0 - s2;
  u64 lo = s0, hi = 0;
  u64 t = s1 << h;
  lo += t;
  hi += (s1 >> (64 - h)) + (lo < t);
  t = s2 << (2 * h);
  lo += t;
  hi += (s2 >> (64 - 2 * h)) + (lo < t);
  return barrett128(lo, hi, m.q, m.ratio_hi, m.ratio_lo);
}

// ---- MAC_INT24: operands < 2^48 split at bit 24; three 64-bit partial sums on the integer pipe ----
template <>
struct Opnd<MAC_INT24> {
  u32 lo, hi, sum;
  __device__ __forceinline__ Opnd(u64 x, int) : lo((u32)x & 0xFFFFFFu), hi((u32)(x >> 24)) { sum = lo + hi; }
};
template <>
struct Acc<MAC_INT24> {
  u64 s0 = 0, sk = 0, s2 = 0;
  __device__ __forceinline__ void mac(const Opnd<MAC_INT24>& a, const Opnd<MAC_INT24>& b) {
    s0 += (u64)a.lo * b.lo;
    sk += (u64)a.sum * b.sum;
    s2 += (u64)a.hi * b.hi;
  }
  __device__ __forceinline__ u64 reduce(const ModC& m, int) const { return karatsuba_reduce(s0, sk, s2, 24, m); }
};

// ---- MAC_FP64: operands < 2^(2h) with h <= 22 split at bit h; the three partial sums live in doubles and every
// product-accumulate is one DFMA on the FP64 pipe.  All values are integers below 2^53, so the arithmetic is exact.
__device__ __forceinline__ double u32_to_double_exact(u32 v) {
  return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0;  // (2^52 + v) - 2^52
}
template <>
struct Opnd<MAC_FP64> {
  double lo, hi, sum;
  __device__ __forceinline__ Opnd(u64 x, int h)
      : lo(u32_to_double_exact((u32)x & ((1u << h) - 1))), hi(u32_to_double_exact((u32)(x >> h))) {
    sum = lo + hi;
  }
};
template <>
struct Acc<MAC_FP64> {
  double s0 = 0.0, sk = 0.0, s2 = 0.0;
  __device__ __forceinline__ void mac(const Opnd<MAC_FP64>& a, const Opnd<MAC_FP64>& b) {
    s0 = fma(a.lo, b.lo, s0);
    sk = fma(a.sum, b.sum, sk);
    s2 = fma(a.hi, b.hi, s2);
  }
  // value = s0 + s1*2^h + s2*2^(2h) mod q entirely on the FP64 pipe: canonicalise the three partial sums, multiply
  // two of them by the precomputed 2^h, 2^(2h) mod q, add, canonicalise.  Returns a canonical integer-valued double.
  __device__ __forceinline__ double reduce_d(const ModC& m) const {
    const double s1 = __dadd_rn(__dadd_rn(sk, -s0), -s2);
    const double r0 = f64_canon(s0, m.qd, m.qinv);
    const double r1 = f64_canon(s1, m.qd, m.qinv);
    const double r2 = f64_canon(s2, m.qd, m.qinv);
    const double t1 = f64_modmul(r1, m.pow_h, m.pow_h_i, m.qd);    // |t| <= 0.54 q
    const double t2 = f64_modmul(r2, m.pow_2h, m.pow_2h_i, m.qd);
    return f64_canon(__dadd_rn(__dadd_rn(r0, t1), t2), m.qd, m.qinv);
  }
  __device__ __forceinline__ u64 reduce(const ModC& m, int) const { return f64_to_u64_exact(reduce_d(m)); }
};

// 128-bit -> [0,q) with ratio = floor(2^128/q)
__device__ __forceinline__ u64 barrett128(u64 lo, u64 hi, u64 q, u64 r_hi, u64 r_lo) {
  u64 carry = __umul64hi(lo, r_lo);
  u64 t_lo = lo * r_hi, t_hi = __umul64hi(lo, r_hi);
  u64 tmp1 = t_lo + carry;
  u64 tmp3 = t_hi + (tmp1 < carry);
  u64 u_lo = hi * r_lo, u_hi = __umul64hi(hi, r_lo);
  u64 tmp1b = tmp1 + u_lo;
  carry = u_hi + (tmp1b < tmp1);
  u64 quot = hi * r_hi + tmp3 + carry;
  return csub(lo - quot * q, q);
}
// 64-bit -> [0,q)
__device__ __forceinline__ u64 barrett64(u64 a, u64 q, u64 r_hi) {
  u64 quot = __umul64hi(a, r_hi);
  return csub(a - quot * q, q);
}
__device__ __forceinline__ u64 mulmod(u64 a, u64 b, const ModC& m) {
  return barrett128(a * b, __umul64hi(a, b), m.q, m.ratio_hi, m.ratio_lo);
}

// ------------------------------------------------------------------------------------------
// Shared-memory layout: XOR swizzle so that every pass of the NTT (element strides
// >=16, 8 and 1) is bank-conflict free for 64-bit words (a half-warp must hit 16
// distinct 8-byte bank pairs).  No padding: a size-N transform uses exactly 8N bytes.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int swz(int i) { return i ^ (((i >> 4) & 7) | (((i >> 6) & 1) << 3)); }

// One pass = R consecutive radix-2 stages done in registers on units of 2^R elements.
// Forward (Cooley-Tukey, SEAL ordering): stage s has m=2^s blocks, gap N>>(s+1), twiddle rp[m+block].
// LAZY (all moduli < 2^58): no per-stage correction at all.  Shoup's lazy product is < 2q for ANY 64-bit input, so
// only the additive path grows, by 2q per stage: after log2(N) stages values are < (4 + 2 log2 N) q <= 32 q < 2^64.
template <int LOGN, int NT, int S0, int R, bool LAZY>
__device__ __forceinline__ void ntt_fwd_pass(u64* __restrict__ s, const u64* __restrict__ rp,
                                             const u64* __restrict__ rps, u64 q, int tid) {
  constexpr int N = 1 << LOGN;
  constexpr int E = 1 << R;
  constexpr int TL = N >> (S0 + R);  // gap of the last stage of this pass
  constexpr int UNITS = N >> R;
  const u64 two_q = 2 * q;
#pragma unroll 1
  for (int u = tid; u < UNITS; u += NT) {
    const int lo = u & (TL - 1);
    const int hi = u / TL;
    const int base = hi * (TL << R) + lo;
    u64 x[E];
#pragma unroll
    for (int e = 0; e < E; ++e) x[e] = s[swz(base + e * TL)];
#pragma unroll
    for (int a = 0; a < R; ++a) {
      const int half = E >> (a + 1);
      const int mbase = (1 << (S0 + a)) + (hi << a);
#pragma unroll
      for (int b = 0; b < (1 << a); ++b) {
        const u64 w = __ldg(rp + mbase + b);
        const u64 ws = __ldg(rps + mbase + b);
#pragma unroll
        for (int c = 0; c < half; ++c) {
          const int e0 = b * 2 * half + c, e1 = e0 + half;
          u64 X = x[e0];
          if (!LAZY) X = X >= two_q ? X - two_q : X;
          const u64 T = shoup_lazy(x[e1], w, ws, q);
          x[e0] = X + T;
          x[e1] = X + two_q - T;
        }
      }
    }
#pragma unroll
    for (int e = 0; e < E; ++e) s[swz(base + e * TL)] = x[e];
  }
}

// Inverse (Gentleman-Sande): same unit geometry, stages visited from the smallest gap up.
// LAZY (all moduli < 2^(62 - log2 N)): sums are never corrected; the bound doubles per stage (B_p = 2^p * 2q for the
// p-th processed stage) and the difference uses X + B_p - Y.  The final N^{-1} Shoup product accepts any 64-bit value.
template <int LOGN, int NT, int S0, int R, bool LAZY>
__device__ __forceinline__ void ntt_inv_pass(u64* __restrict__ s, const u64* __restrict__ irp,
                                             const u64* __restrict__ irps, u64 q, int tid) {
  constexpr int N = 1 << LOGN;
  constexpr int E = 1 << R;
  constexpr int TL = N >> (S0 + R);
  constexpr int UNITS = N >> R;
  const u64 two_q = 2 * q;
#pragma unroll 1
  for (int u = tid; u < UNITS; u += NT) {
    const int lo = u & (TL - 1);
    const int hi = u / TL;
    const int base = hi * (TL << R) + lo;
    u64 x[E];
#pragma unroll
    for (int e = 0; e < E; ++e) x[e] = s[swz(base + e * TL)];
#pragma unroll
    for (int a = R - 1; a >= 0; --a) {
      const int half = E >> (a + 1);
      const int mbase = (1 << (S0 + a)) + (hi << a);
#pragma unroll
      for (int b = 0; b < (1 << a); ++b) {
        const u64 w = __ldg(irp + mbase + b);
        const u64 ws = __ldg(irps + mbase + b);
#pragma unroll
        for (int c = 0; c < half; ++c) {
          const int e0 = b * 2 * half + c, e1 = e0 + half;
          const u64 X = x[e0], Y = x[e1];
          u64 S = X + Y;
          if (!LAZY) S = S >= two_q ? S - two_q : S;
          const u64 D = LAZY ? X + (two_q << (LOGN - 1 - (S0 + a))) - Y : X + two_q - Y;
          x[e0] = S;
          x[e1] = shoup_lazy(D, w, ws, q);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < E; ++e) s[swz(base + e * TL)] = x[e];
  }
}

// Full transforms on a swizzled shared-memory buffer.  Stage split: first pass takes
// R0 = ((LOGN-1) % 3) + 1 stages, every later pass 3 stages, so the last-stage gaps of the
// passes are always ..., 64, 8, 1 (the cases the swizzle is designed for).
// Input of forward: values < 4q.  Output: values < 4q (caller reduces).
template <int LOGN, int NT, bool LAZY>
__device__ __forceinline__ void ntt_forward_smem_t(u64* s, const ModC& m, int tid) {
  constexpr int R0 = ((LOGN - 1) % 3) + 1;
  const u64 q = m.q;
  ntt_fwd_pass<LOGN, NT, 0, R0, LAZY>(s, m.rp, m.rps, q, tid);
  __syncthreads();
  if constexpr (LOGN > R0) {
    ntt_fwd_pass<LOGN, NT, R0, 3, LAZY>(s, m.rp, m.rps, q, tid);
    __syncthreads();
  }
  if constexpr (LOGN > R0 + 3) {
    ntt_fwd_pass<LOGN, NT, R0 + 3, 3, LAZY>(s, m.rp, m.rps, q, tid);
    __syncthreads();
  }
  if constexpr (LOGN > R0 + 6) {
    ntt_fwd_pass<LOGN, NT, R0 + 6, 3, LAZY>(s, m.rp, m.rps, q, tid);
    __syncthreads();
  }
  if constexpr (LOGN > R0 + 9) {
    ntt_fwd_pass<LOGN, NT, R0 + 9, 3, LAZY>(s, m.rp, m.rps, q, tid);
    __syncthreads();
  }
  static_assert(LOGN <= R0 + 12, "unsupported transform size");
}
// Input of inverse: values < 2q.  Output: values < 2q, NOT yet scaled by N^{-1}.
template <int LOGN, int NT, bool LAZY>
__device__ __forceinline__ void ntt_inverse_smem_t(u64* s, const ModC& m, int tid) {
  constexpr int R0 = ((LOGN - 1) % 3) + 1;
  const u64 q = m.q;
  if constexpr (LOGN > R0 + 9) {
    ntt_inv_pass<LOGN, NT, R0 + 9, 3, LAZY>(s, m.irp, m.irps, q, tid);
    __syncthreads();
  }
  if constexpr (LOGN > R0 + 6) {
    ntt_inv_pass<LOGN, NT, R0 + 6, 3, LAZY>(s, m.irp, m.irps, q, tid);
    __syncthreads();
  }
  if constexpr (LOGN > R0 + 3) {
    ntt_inv_pass<LOGN, NT, R0 + 3, 3, LAZY>(s, m.irp, m.irps, q, tid);
    __syncthreads();
  }
  if constexpr (LOGN > R0) {
    ntt_inv_pass<LOGN, NT, R0, 3, LAZY>(s, m.irp, m.irps, q, tid);
    __syncthreads();
  }
  ntt_inv_pass<LOGN, NT, 0, R0, LAZY>(s, m.irp, m.irps, q, tid);
  __syncthreads();
}

// reduce a forward-NTT output of the integer engines to canonical: < 4q (corrected butterflies) or < 32q (lazy)
__device__ __forceinline__ u64 canon_fwd(u64 v, const ModC& m, bool lazy) {
  if (lazy) return barrett64(v, m.q, m.ratio_hi);
  v = v >= 2 * m.q ? v - 2 * m.q : v;
  return csub(v, m.q);
}
// scale an inverse-NTT lazy value by N^{-1} and make canonical (accepts any 64-bit input)
__device__ __forceinline__ u64 inv_finish(u64 v, const ModC& m) { return shoup(v, m.inv_n, m.inv_n_s, m.q); }

// ------------------------------------------------------------------------------------------
// FP64 engine (moduli <= 44 bits): the polynomial lives in shared memory as integer-valued doubles and a
// butterfly is 8 FP64-pipe instructions.  Exact modular product of an integer-valued double y (|y| < 2^48) with a
// table constant w in [0,q) (wi = w/q rounded):
//     h = y*w (rounded)          l = fma(y, w, -h)      (exact low part, |l| <= ulp(h)/2)
//     c = rint(y*wi)             via fma(y, wi, 1.5*2^52) - 1.5*2^52
//     r = fma(-c, q, h)          (exact: an integer below 2^46)        t = r + l = y*w - c*q,  |t| <= 0.54 q
// Every intermediate is an integer below 2^53, so results are exact; values grow by <= 0.54 q per stage
// (< 9q after 14 stages from a canonical input).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double f64_modmul(double y, double w, double wi, double q) {
  const double M = 6755399441055744.0;  // 1.5 * 2^52
  const double h = __dmul_rn(y, w);
  const double l = __fma_rn(y, w, -h);
  const double c = __dadd_rn(__fma_rn(y, wi, M), -M);
  const double r = __fma_rn(-c, q, h);
  return __dadd_rn(r, l);
}
// integer-valued double v (|v| < 2^51) -> canonical residue as a double in [0,q)
__device__ __forceinline__ double f64_canon(double v, double q, double qinv) {
  const double M = 6755399441055744.0;
  const double c = __dadd_rn(__fma_rn(v, qinv, M), -M);
  double r = __fma_rn(-c, q, v);
  return r < 0.0 ? __dadd_rn(r, q) : r;
}
__device__ __forceinline__ double u64_to_f64_exact(u64 v) {  // v < 2^52
  return __dadd_rn(__longlong_as_double((long long)(v | 0x4330000000000000ull)), -4503599627370496.0);
}
__device__ __forceinline__ u64 f64_to_u64_exact(double d) {  // integer-valued, 0 <= d < 2^52
  return (u64)__double_as_longlong(__dadd_rn(d, 4503599627370496.0)) & 0x000FFFFFFFFFFFFFull;
}

// forward pass (Cooley-Tukey, SEAL ordering) on doubles; same unit geometry as ntt_fwd_pass
// Twiddle sources of the FP64 passes: the (w, w/q) pair table in global memory, or a w-only table staged in shared
// memory whose companion is formed on the fly (w * (1/q): relative error 2^-52, so the quotient estimate of
// f64_modmul stays within 0.5 + |y| 2^-51 of the true quotient and every bound above holds with 0.57 q per stage).
struct TwGlobal {
  const double2* t;
  __device__ __forceinline__ void get(int i, double& w, double& wi) const {
    const double2 v = __ldg(t + i);
    w = v.x;
    wi = v.y;
  }
};
struct TwShared {
  const double* t;
  double qinv;
  __device__ __forceinline__ void get(int i, double& w, double& wi) const {
    w = t[i];
    wi = __dmul_rn(w, qinv);
  }
};

template <int LOGN, int NT, int S0, int R, class TW>
__device__ __forceinline__ void f64_fwd_pass(double* __restrict__ s, const TW tw, double q, int tid) {
  constexpr int N = 1 << LOGN;
  constexpr int E = 1 << R;
  constexpr int TL = N >> (S0 + R);
  constexpr int UNITS = N >> R;
#pragma unroll 1
  for (int u = tid; u < UNITS; u += NT) {
    const int lo = u & (TL - 1);
    const int hi = u / TL;
    const int base = hi * (TL << R) + lo;
    double x[E];
#pragma unroll
    for (int e = 0; e < E; ++e) x[e] = s[swz(base + e * TL)];
#pragma unroll
    for (int a = 0; a < R; ++a) {
      const int half = E >> (a + 1);
      const int mbase = (1 << (S0 + a)) + (hi << a);
#pragma unroll
      for (int b = 0; b < (1 << a); ++b) {
        double w, wi;
        tw.get(mbase + b, w, wi);
#pragma unroll
        for (int c = 0; c < half; ++c) {
          const int e0 = b * 2 * half + c, e1 = e0 + half;
          const double T = f64_modmul(x[e1], w, wi, q);
          x[e1] = __dadd_rn(x[e0], -T);
          x[e0] = __dadd_rn(x[e0], T);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < E; ++e) s[swz(base + e * TL)] = x[e];
  }
}
// inverse pass: decimation-in-time butterflies (X + W*Y, X - W*Y) on the bit-reversed input, gaps increasing.
// For the stage with gap g the twiddle of the pair at in-block offset j is iw[g + j] = psi^(-j*N/g)  (cyclic
// inverse DFT with root psi^-2); the remaining psi^-i * N^-1 is applied per element at the end (fin table).
template <int LOGN, int NT, int S0, int R, class TW>
__device__ __forceinline__ void f64_inv_pass(double* __restrict__ s, const TW iw, double q, int tid) {
  constexpr int N = 1 << LOGN;
  constexpr int E = 1 << R;
  constexpr int TL = N >> (S0 + R);
  constexpr int UNITS = N >> R;
#pragma unroll 1
  for (int u = tid; u < UNITS; u += NT) {
    const int lo = u & (TL - 1);
    const int hi = u / TL;
    const int base = hi * (TL << R) + lo;
    double x[E];
#pragma unroll
    for (int e = 0; e < E; ++e) x[e] = s[swz(base + e * TL)];
#pragma unroll
    for (int a = R - 1; a >= 0; --a) {
      const int dist = E >> (a + 1);   // pair distance in unit elements; gap = dist * TL
      const int g = dist * TL;
#pragma unroll
      for (int c = 0; c < dist; ++c) {  // in-block offset j = c*TL + lo
        double w, wi;
        iw.get(g + c * TL + lo, w, wi);
#pragma unroll
        for (int b = 0; b < (1 << a); ++b) {
          const int e0 = b * 2 * dist + c, e1 = e0 + dist;
          const double T = f64_modmul(x[e1], w, wi, q);
          x[e1] = __dadd_rn(x[e0], -T);
          x[e0] = __dadd_rn(x[e0], T);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < E; ++e) s[swz(base + e * TL)] = x[e];
  }
}

// ------------------------------------------------------------------------------------------
// Engine front-end used by every NTT-based kernel.  ENG: 0 integer (corrected butterflies, any modulus < 2^61),
// 1 integer fully lazy, 2 FP64.  Shared memory holds one 64-bit word per coefficient in all engines.
// ------------------------------------------------------------------------------------------
enum { ENG_INT = 0, ENG_INT_LAZY = 1, ENG_FP64 = 2 };

template <int ENG>
__device__ __forceinline__ u64 eng_load(u64 v) {  // canonical (or < 4q) value -> shared-memory word
  if constexpr (ENG == ENG_FP64) return (u64)__double_as_longlong(u64_to_f64_exact(v));
  else return v;
}
template <int ENG>
__device__ __forceinline__ u64 eng_store_fwd(u64 word, const ModC& m) {  // forward output word -> canonical
  if constexpr (ENG == ENG_FP64) return f64_to_u64_exact(f64_canon(__longlong_as_double((long long)word), m.qd, m.qinv));
  else return canon_fwd(word, m, ENG == ENG_INT_LAZY);
}
template <int ENG>
__device__ __forceinline__ u64 eng_store_inv(u64 word, int i, const ModC& m) {  // inverse output word -> canonical
  if constexpr (ENG == ENG_FP64) {
    const double2 f2 = __ldg(m.fin + i);
    double t = f64_modmul(__longlong_as_double((long long)word), f2.x, f2.y, m.qd);
    t = t < 0.0 ? __dadd_rn(t, m.qd) : t;
    return f64_to_u64_exact(t);
  } else {
    return inv_finish(word, m);
  }
}
// canonical doubles in [0,q)
__device__ __forceinline__ double f64_addmod(double a, double b, double q) {
  const double s = __dadd_rn(a, b);
  return s >= q ? __dadd_rn(s, -q) : s;
}
__device__ __forceinline__ double f64_submod(double a, double b, double q) {
  const double d = __dadd_rn(a, -b);
  return d < 0.0 ? __dadd_rn(d, q) : d;
}
__device__ __forceinline__ double f64_negmod(double a, double q) { return a == 0.0 ? 0.0 : __dadd_rn(q, -a); }
// inverse output word -> canonical value kept in the engine's own representation (double bits for ENG_FP64)
template <int ENG>
__device__ __forceinline__ u64 eng_finish_inv_native(u64 word, int i, const ModC& m) {
  if constexpr (ENG == ENG_FP64) {
    const double2 f2 = __ldg(m.fin + i);
    double t = f64_modmul(__longlong_as_double((long long)word), f2.x, f2.y, m.qd);
    t = t < 0.0 ? __dadd_rn(t, m.qd) : t;
    return (u64)__double_as_longlong(t);
  } else {
    return inv_finish(word, m);
  }
}

template <int LOGN, int NT, class TW>
__device__ __forceinline__ void f64_forward_all(double* d, const TW tw, double q, int tid) {
  constexpr int R0 = ((LOGN - 1) % 3) + 1;
  f64_fwd_pass<LOGN, NT, 0, R0>(d, tw, q, tid);
  __syncthreads();
  if constexpr (LOGN > R0) { f64_fwd_pass<LOGN, NT, R0, 3>(d, tw, q, tid); __syncthreads(); }
  if constexpr (LOGN > R0 + 3) { f64_fwd_pass<LOGN, NT, R0 + 3, 3>(d, tw, q, tid); __syncthreads(); }
  if constexpr (LOGN > R0 + 6) { f64_fwd_pass<LOGN, NT, R0 + 6, 3>(d, tw, q, tid); __syncthreads(); }
  if constexpr (LOGN > R0 + 9) { f64_fwd_pass<LOGN, NT, R0 + 9, 3>(d, tw, q, tid); __syncthreads(); }
}
template <int LOGN, int NT, class TW>
__device__ __forceinline__ void f64_inverse_all(double* d, const TW iw, double q, int tid) {
  constexpr int R0 = ((LOGN - 1) % 3) + 1;
  if constexpr (LOGN > R0 + 9) { f64_inv_pass<LOGN, NT, R0 + 9, 3>(d, iw, q, tid); __syncthreads(); }
  if constexpr (LOGN > R0 + 6) { f64_inv_pass<LOGN, NT, R0 + 6, 3>(d, iw, q, tid); __syncthreads(); }
  if constexpr (LOGN > R0 + 3) { f64_inv_pass<LOGN, NT, R0 + 3, 3>(d, iw, q, tid); __syncthreads(); }
  if constexpr (LOGN > R0) { f64_inv_pass<LOGN, NT, R0, 3>(d, iw, q, tid); __syncthreads(); }
  f64_inv_pass<LOGN, NT, 0, R0>(d, iw, q, tid);
  __syncthreads();
}
template <int LOGN, int NT, int ENG>
__device__ __forceinline__ void eng_forward(u64* s, const ModC& m, int tid) {
  if constexpr (ENG == ENG_FP64) f64_forward_all<LOGN, NT>(reinterpret_cast<double*>(s), TwGlobal{m.fw}, m.qd, tid);
  else ntt_forward_smem_t<LOGN, NT, ENG == ENG_INT_LAZY>(s, m, tid);
}
template <int LOGN, int NT, int ENG>
__device__ __forceinline__ void eng_inverse(u64* s, const ModC& m, int tid) {
  if constexpr (ENG == ENG_FP64) f64_inverse_all<LOGN, NT>(reinterpret_cast<double*>(s), TwGlobal{m.iw}, m.qd, tid);
  else ntt_inverse_smem_t<LOGN, NT, ENG == ENG_INT_LAZY>(s, m, tid);
}

// ------------------------------------------------------------------------------------------
// mbarrier + bulk asynchronous copy (cp.async.bulk, the non-tensor TMA path) helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u64* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, u32 bytes, u64* bar, u64 policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

__device__ __forceinline__ void bulk_g2s_plain(void* dst, const void* src, u32 bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Galois automorphism x -> x^g in coefficient form, as a gather: value of sigma_g(a) at index n.
// ginv = g^{-1} mod 2N.  (out[i*g mod N] = +-in[i], sign from (i*g div N) parity.)
__device__ __forceinline__ u64 galois_gather(const u64* __restrict__ in, u32 n, u32 ginv, u32 N, u64 q) {
  u32 i1 = (n * ginv) & (2 * N - 1);
  if (i1 < N) return in[i1];
  return negmod(in[i1 - N], q);
}

}  // namespace pirb
