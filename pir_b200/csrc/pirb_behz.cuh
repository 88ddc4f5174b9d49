// pirb_behz.cuh — per-coefficient arithmetic of the ciphertext-multiplication mode (database.cpp:202-211):
// Evaluator::multiply in SEAL 3.5.6 is the BEHZ RNS variant of BFV multiplication (util/rns.cpp RNSTool,
// evaluator.cpp bfv_multiply steps (1)-(8)).  Everything outside the NTTs and the dyadic products is independent per
// coefficient, so each step is a function of one coefficient's residues; the kernels in kernels_ctmul.cu only add the
// indexing.  Plain integer code on purpose: the same functions compile for the host and are checked there against the
// oracle (tests/cpp/device_math_host_test.cpp), where no GPU is available.
//
// Bases: q = q_0..q_{k-1} (data level), B = b_0..b_{nB-1}, Bsk = B u {m_sk} (m_sk last), m_tilde = 2^32.
#pragma once
#include <cuda_runtime.h>

#include "pirb_common.h"

namespace pirb {

__device__ __forceinline__ u64 lite_csub(u64 x, u64 q) { return x >= q ? x - q : x; }
// 128-bit -> [0,q) with ratio = floor(2^128/q)  (same reduction as barrett128 in pirb_device.cuh)
__device__ __forceinline__ u64 lite_barrett128(u64 lo, u64 hi, const ModLite& m) {
  u64 carry = __umul64hi(lo, m.ratio_lo);
  u64 t_lo = lo * m.ratio_hi, t_hi = __umul64hi(lo, m.ratio_hi);
  u64 tmp1 = t_lo + carry;
  u64 tmp3 = t_hi + (tmp1 < carry);
  u64 u_lo = hi * m.ratio_lo, u_hi = __umul64hi(hi, m.ratio_lo);
  u64 tmp1b = tmp1 + u_lo;
  carry = u_hi + (tmp1b < tmp1);
  u64 quot = hi * m.ratio_hi + tmp3 + carry;
  return lite_csub(lo - quot * m.q, m.q);
}
__device__ __forceinline__ u64 lite_mulmod(u64 a, u64 b, const ModLite& m) {
  return lite_barrett128(a * b, __umul64hi(a, b), m);
}
// (hi:lo) += a*b without inline PTX (host-compilable); at most PIRB_MAX_BSK products below 2^123 are ever summed
__device__ __forceinline__ void lite_mac(u64& lo, u64& hi, u64 a, u64 b) {
  const u64 pl = a * b, ph = __umul64hi(a, b);
  lo += pl;
  hi += ph + (lo < pl);
}
__device__ __forceinline__ void lite_add(u64& lo, u64& hi, u64 v) {
  lo += v;
  hi += (lo < v);
}

// steps (1)-(2) for one coefficient: x[j] (residues mod q_j, canonical) -> y[i] (residues mod bsk_i, canonical)
//   (1) fastbconv_m_tilde: FastBConv_{q -> Bsk u {m_tilde}}(x * m_tilde)
//   (2) sm_mrq: r = -(that mod m_tilde) / Q mod m_tilde, centred; y_i = (conv_i + Q r) / m_tilde mod bsk_i
__device__ __forceinline__ void behz_extend_coeff(const BehzC& B, const u64* x, u64* y) {
  u64 tmp[PIRB_MAX_MODULI];
  u64 r_mt = 0;
  for (int j = 0; j < B.k; ++j) {
    tmp[j] = lite_mulmod(lite_mulmod(x[j], B.mtilde_mod_q[j], B.q[j]), B.inv_qhat_mod_q[j], B.q[j]);
    r_mt += tmp[j] * B.qhat_mod_mtilde[j];  // modulo 2^64, and 2^32 divides 2^64
  }
  const u64 mask = 0xFFFFFFFFull;
  r_mt = ((r_mt & mask) * B.neg_inv_q_mod_mtilde) & mask;
  for (int i = 0; i <= B.nB; ++i) {
    const ModLite& p = B.bsk[i];
    u64 lo = 0, hi = 0;
    for (int j = 0; j < B.k; ++j) lite_mac(lo, hi, tmp[j], B.qhat_mod_bsk[i][j]);
    const u64 conv = lite_barrett128(lo, hi, p);
    u64 r = r_mt;
    if (r >= 0x80000000ull) r += p.q - 0x100000000ull;
    lo = 0;
    hi = 0;
    lite_mac(lo, hi, r, B.q_mod_bsk[i]);
    lite_add(lo, hi, conv);
    y[i] = lite_mulmod(lite_barrett128(lo, hi, p), B.inv_mtilde_mod_bsk[i], p);
  }
}

// steps (6)-(8) for one coefficient: dq[j] mod q_j, db[i] mod bsk_i (the tensor product, coefficient form) -> out[j]
//   (6) multiply by t   (7) fast_floor: divide by Q and floor, result in Bsk
//   (8) fastbconv_sk: back to q with the Shenoy-Kumaresan correction taken from the m_sk residue
__device__ __forceinline__ void behz_floor_coeff(const BehzC& B, const u64* dq, const u64* db, u64* out) {
  u64 u[PIRB_MAX_MODULI], f[PIRB_MAX_BSK], v[PIRB_MAX_BSK];
  for (int j = 0; j < B.k; ++j)
    u[j] = lite_mulmod(lite_mulmod(dq[j], B.t_mod_q[j], B.q[j]), B.inv_qhat_mod_q[j], B.q[j]);
  for (int i = 0; i <= B.nB; ++i) {
    const ModLite& p = B.bsk[i];
    u64 lo = 0, hi = 0;
    for (int j = 0; j < B.k; ++j) lite_mac(lo, hi, u[j], B.qhat_mod_bsk[i][j]);
    const u64 conv = lite_barrett128(lo, hi, p);
    const u64 tb = lite_mulmod(db[i], B.t_mod_bsk[i], p);
    f[i] = lite_mulmod(tb + (p.q - conv), B.inv_q_mod_bsk[i], p);
  }
  const ModLite& msk = B.bsk[B.nB];
  u64 lo = 0, hi = 0;
  for (int j = 0; j < B.nB; ++j) {
    v[j] = lite_mulmod(f[j], B.inv_bhat_mod_b[j], B.bsk[j]);
    lite_mac(lo, hi, v[j], B.bhat_mod_msk[j]);
  }
  const u64 a_sk = lite_barrett128(lo, hi, msk);
  const u64 alpha = lite_mulmod(a_sk + (msk.q - f[B.nB]), B.inv_b_mod_msk, msk);
  const bool negative = alpha > (msk.q >> 1);
  for (int i = 0; i < B.k; ++i) {
    const ModLite& m = B.q[i];
    lo = 0;
    hi = 0;
    for (int j = 0; j < B.nB; ++j) lite_mac(lo, hi, v[j], B.bhat_mod_q[i][j]);
    const u64 conv = lite_barrett128(lo, hi, m);
    lo = 0;
    hi = 0;
    if (negative) lite_mac(lo, hi, msk.q - alpha, B.b_mod_q[i]);
    else lite_mac(lo, hi, alpha, m.q - B.b_mod_q[i]);
    lite_add(lo, hi, conv);
    out[i] = lite_barrett128(lo, hi, m);
  }
}

// step (4) for one coefficient and one modulus: D_i = sum_{x+y=i} A_x B_y with B of two polynomials.
// a[0..s1), b0, b1 canonical; d[0..s1] canonical.
__device__ __forceinline__ void behz_tensor_coeff(const ModLite& m, const u64* a, int s1, u64 b0, u64 b1, u64* d) {
  d[0] = lite_mulmod(a[0], b0, m);
  for (int i = 1; i < s1; ++i) {
    u64 lo = 0, hi = 0;
    lite_mac(lo, hi, a[i], b0);
    lite_mac(lo, hi, a[i - 1], b1);
    d[i] = lite_barrett128(lo, hi, m);
  }
  d[s1] = lite_mulmod(a[s1 - 1], b1, m);
}

// Relinearization (Evaluator::relinearize_internal -> switch_key_inplace, BFV branch), per coefficient.
// mod-down of one accumulator pair by the special prime with rounding: acc_j (mod q_j), last (mod P), all in
// coefficient form -> the value to add to the ciphertext component mod q_j
__device__ __forceinline__ u64 relin_moddown_coeff(u64 acc_j, u64 last, u64 P, u64 half_P, u64 half_P_mod_qj, u64 inv_P_mod_qj,
                                                   const ModLite& qj) {
  u64 l = last + half_P;
  l = l >= P ? l - P : l;
  u64 r = lite_barrett128(l, 0, qj);
  r = r >= half_P_mod_qj ? r - half_P_mod_qj : r + qj.q - half_P_mod_qj;
  const u64 diff = acc_j >= r ? acc_j - r : acc_j + qj.q - r;
  return lite_mulmod(diff, inv_P_mod_qj, qj);
}

// ------------------------------------------------------------------------------------------
// Kernel bodies of kernels_ctmul.cu: one call = one thread (idx = global thread index along x, y_ = batch index).
// ------------------------------------------------------------------------------------------
// steps (1)-(2): in [p][k][N] base q, coefficient form -> out [p][nB+1][N] base Bsk, coefficient form
__device__ __forceinline__ void k_behz_extend_body(const BehzC& B, u64 idx, u64 y_, const u64* __restrict__ in,
    u64 in_bstride, u32 n_polys, u64* __restrict__ out, u64 out_bstride) {
  const u32 N = B.N;
  if (idx >= (u64)n_polys * N) return;
  const u32 n = (u32)idx & (N - 1);
  const u64 p = idx >> B.logn;
  const u64* src = in + y_ * in_bstride + p * B.k * N + n;
  u64 x[PIRB_MAX_DATA], y[PIRB_MAX_BSK];
  for (int j = 0; j < B.k; ++j) x[j] = src[(u64)j * N];
  behz_extend_coeff(B, x, y);
  u64* dst = out + y_ * out_bstride + p * (B.nB + 1) * N + n;
  for (int i = 0; i <= B.nB; ++i) dst[(u64)i * N] = y[i];
}

// step (4) in one base (base = 0: q, 1: Bsk), NTT form:
//   A [e][s1][nm][N], S [i][2][nm][N], D [e][s1+1][nm][N];  entry e multiplies selection entry e % dim
__device__ __forceinline__ void k_behz_tensor_body(const BehzC& B, u64 idx, u64 y_, int base,
    const u64* __restrict__ A, u64 a_bstride, const u64* __restrict__ S, u64 s_bstride, u64* __restrict__ D,
    u64 d_bstride, u32 n_entries, u32 dim, int s1) {
  const u32 N = B.N;
  const u32 nm = base ? (u32)B.nB + 1 : (u32)B.k;
  if (idx >= (u64)n_entries * nm * N) return;
  const u32 n = (u32)idx & (N - 1);
  const u64 pm = idx >> B.logn;
  const u32 jm = (u32)(pm % nm);
  const u64 e = pm / nm;
  const ModLite& m = base ? B.bsk[jm] : B.q[jm];
  const u64* a = A + y_ * a_bstride + ((e * s1) * nm + jm) * N + n;
  const u64* s = S + y_ * s_bstride + (((e % dim) * 2) * nm + jm) * N + n;
  u64 av[PIRB_MAX_DIMS + 2], dv[PIRB_MAX_DIMS + 3];
  for (int x = 0; x < s1; ++x) av[x] = a[(u64)x * nm * N];
  behz_tensor_coeff(m, av, s1, s[0], s[(u64)nm * N], dv);
  u64* d = D + y_ * d_bstride + ((e * (s1 + 1)) * nm + jm) * N + n;
  for (int x = 0; x <= s1; ++x) d[(u64)x * nm * N] = dv[x];
}

// steps (6)-(8): Dq [p][k][N], Db [p][nB+1][N] (coefficient form) -> out [p][k][N]
__device__ __forceinline__ void k_behz_floor_body(const BehzC& B, u64 idx, u64 y_, const u64* __restrict__ Dq,
    u64 dq_bstride, const u64* __restrict__ Db, u64 db_bstride, u64* __restrict__ out, u64 out_bstride,
    u32 n_polys) {
  const u32 N = B.N;
  if (idx >= (u64)n_polys * N) return;
  const u32 n = (u32)idx & (N - 1);
  const u64 p = idx >> B.logn;
  const u64* sq = Dq + y_ * dq_bstride + p * B.k * N + n;
  const u64* sb = Db + y_ * db_bstride + p * (B.nB + 1) * N + n;
  u64 dq[PIRB_MAX_DATA], db[PIRB_MAX_BSK], o[PIRB_MAX_DATA];
  for (int j = 0; j < B.k; ++j) dq[j] = sq[(u64)j * N];
  for (int i = 0; i <= B.nB; ++i) db[i] = sb[(u64)i * N];
  behz_floor_coeff(B, dq, db, o);
  u64* dst = out + y_ * out_bstride + p * B.k * N + n;
  for (int j = 0; j < B.k; ++j) dst[(u64)j * N] = o[j];
}

// switch_key_inplace, step 1: third polynomial of product e (prod [e][3][k][N]) -> dig [e][J][I][N], digit J re-reduced
// modulo key-level modulus I (I = k: the special prime); the forward NTT follows (launch_ntt_fwd, cycle k + 1)
__device__ __forceinline__ void k_relin_digits_body(const BehzC& B, u64 idx, u64 y_, const u64* __restrict__ prod,
    u64 p_bstride, u64* __restrict__ dig, u64 dig_bstride, u32 n_entries) {
  const u32 N = B.N;
  const u32 k = (u32)B.k, k1 = k + 1;
  if (idx >= (u64)n_entries * k * k1 * N) return;
  const u32 n = (u32)idx & (N - 1);
  const u64 pm = idx >> B.logn;
  const u32 I = (u32)(pm % k1);
  const u32 J = (u32)((pm / k1) % k);
  const u64 e = pm / ((u64)k1 * k);
  const ModLite& mI = I < k ? B.q[I] : B.P;
  u64 v = prod[y_ * p_bstride + ((e * 3 + 2) * k + J) * N + n];
  if (B.q[J].q > mI.q) v = lite_barrett128(v, 0, mI);
  dig[y_ * dig_bstride + ((e * k + J) * k1 + I) * N + n] = v;
}

// step 2: acc [e][c][I][N] = sum_J dig[e][J][I] (.) key[J][c][I]   (NTT form; the inverse NTT follows)
__device__ __forceinline__ void k_relin_mac_body(const BehzC& B, u64 idx, u64 y_, const u64* __restrict__ dig,
    u64 dig_bstride, const u64* __restrict__ key, u64* __restrict__ acc, u64 acc_bstride, u32 n_entries) {
  const u32 N = B.N;
  const u32 k = (u32)B.k, k1 = k + 1;
  if (idx >= (u64)n_entries * 2 * k1 * N) return;
  const u32 n = (u32)idx & (N - 1);
  const u64 pm = idx >> B.logn;
  const u32 I = (u32)(pm % k1);
  const u32 c = (u32)((pm / k1) & 1);
  const u64 e = pm / (2ull * k1);
  const ModLite& mI = I < k ? B.q[I] : B.P;
  const u64* d = dig + y_ * dig_bstride + ((e * k) * k1 + I) * N + n;
  u64 lo = 0, hi = 0;
  for (u32 J = 0; J < k; ++J) lite_mac(lo, hi, d[(u64)J * k1 * N], __ldg(key + ((u64)(J * 2 + c) * k1 + I) * N + n));
  acc[y_ * acc_bstride + ((e * 2 + c) * k1 + I) * N + n] = lite_barrett128(lo, hi, mI);
}

// step 3: mod-down by the special prime with rounding, added to the first two polynomials of the product:
//   X [e][c][j][N] = prod[e][c][j] + (acc[e][c][j] - round-term(acc[e][c][k])) / P      (acc in coefficient form)
__device__ __forceinline__ void k_relin_finish_body(const BehzC& B, u64 idx, u64 y_, const u64* __restrict__ prod,
    u64 p_bstride, const u64* __restrict__ acc, u64 acc_bstride, u64* __restrict__ X, u64 x_bstride, u32 n_entries) {
  const u32 N = B.N;
  const u32 k = (u32)B.k, k1 = k + 1;
  if (idx >= (u64)n_entries * 2 * k * N) return;
  const u32 n = (u32)idx & (N - 1);
  const u64 pm = idx >> B.logn;
  const u32 j = (u32)(pm % k);
  const u32 c = (u32)((pm / k) & 1);
  const u64 e = pm / (2ull * k);
  const u64* a = acc + y_ * acc_bstride + ((e * 2 + c) * k1) * N + n;
  const u64 delta = relin_moddown_coeff(a[(u64)j * N], a[(u64)k * N], B.P.q, B.half_P, B.half_P_mod_q[j], B.inv_P_mod_q[j],
                                        B.q[j]);
  const u64 v = prod[y_ * p_bstride + ((e * 3 + c) * k + j) * N + n] + delta;
  X[y_ * x_bstride + ((e * 2 + c) * k + j) * N + n] = lite_csub(v, B.q[j].q);
}

// database.cpp:240-247: out [g][polys][k][N] = sum_{i < cnt(g)} X[g*dim + i][polys][k][N]  (mod q_j), cnt(g) = entries left
__device__ __forceinline__ void k_ct_reduce_body(const BehzC& B, u64 idx, u64 y_, const u64* __restrict__ X,
    u64 x_bstride, u64* __restrict__ out, u64 out_bstride, u32 n_entries, u32 dim, u32 polys) {
  const u32 N = B.N;
  const u32 k = (u32)B.k;
  const u32 n_groups = (n_entries + dim - 1) / dim;
  const u64 per_ct = (u64)polys * k * N;
  if (idx >= (u64)n_groups * per_ct) return;
  const u64 g = idx / per_ct;
  const u64 within = idx - g * per_ct;
  const u32 j = (u32)((within >> B.logn) % k);
  const u64 q = B.q[j].q;
  const u32 first = (u32)g * dim;
  const u32 cnt = n_entries - first < dim ? n_entries - first : dim;
  const u64* src = X + y_ * x_bstride + (u64)first * per_ct + within;
  u64 v = src[0];
  for (u32 i = 1; i < cnt; ++i) v = lite_csub(v + src[(u64)i * per_ct], q);
  out[y_ * out_bstride + idx] = v;
}


}  // namespace pirb
