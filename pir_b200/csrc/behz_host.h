// behz_host.h — host-side setup of the ciphertext-multiplication mode (product code; independent of oracle/).
// Restates what SEAL 3.5.6 computes in RNSTool::initialize (util/rns.cpp) for the first data level: the auxiliary
// primes from get_primes(N, 61, |B| + 2) (util/numth.cpp: 2^61 - i*2N + 1, descending; the first is m_sk, the second
// gamma — used only by decryption —, the rest B), |B| = |q| (+1 if m_tilde*Q would not fit), m_tilde = 2^32, and the
// base-conversion constants of fastbconv_m_tilde / sm_mrq / fast_floor / fastbconv_sk.
#pragma once
#include <vector>

#include "host_math.h"
#include "pirb_common.h"

namespace pirb {
namespace hm {

inline std::vector<u64> get_primes(u64 N, int bit_size, size_t count) {
  std::vector<u64> out;
  const u64 factor = 2 * N;
  u64 value = (1ull << bit_size) - factor + 1;
  const u64 lower = 1ull << (bit_size - 1);
  while (count > 0 && value > lower) {
    if (is_prime(value)) { out.push_back(value); --count; }
    value -= factor;
  }
  return out;  // shorter than asked for: the caller reports the error
}

inline int product_bits(const u64* primes, int n) {
  std::vector<u64> v{1};
  for (int i = 0; i < n; ++i) {
    u128 carry = 0;
    for (auto& limb : v) { carry += (u128)limb * primes[i]; limb = (u64)carry; carry >>= 64; }
    if (carry) v.push_back((u64)carry);
  }
  while (v.size() > 1 && !v.back()) v.pop_back();
  return (int)(v.size() - 1) * 64 + (64 - __builtin_clzll(v.back()));
}

// product of primes[0..n) except index `skip` (skip < 0: none), modulo p (any modulus, p > 1)
inline u64 prod_mod(const u64* primes, int n, int skip, u64 p) {
  u64 r = 1 % p;
  for (int i = 0; i < n; ++i)
    if (i != skip) r = mulmod(r, primes[i] % p, p);
  return r;
}

inline ModLite mod_lite(u64 q) {
  ModLite m;
  m.q = q;
  barrett_ratio(q, &m.ratio_hi, &m.ratio_lo);
  return m;
}

// q[0..k): data moduli, P: special prime, t: plain modulus.  bsk_out receives the nB + 1 primes of Bsk (m_sk last).
// Returns false if the auxiliary primes cannot be found or collide with the coefficient moduli.
inline bool build_behz(const u64* q, int k, u64 P, u32 N, int logn, u64 t, BehzC* out, std::vector<u64>* bsk_out) {
  if (k < 1 || k > PIRB_MAX_DATA) return false;
  BehzC& C = *out;
  C = BehzC();
  C.k = k;
  C.N = N;
  C.logn = logn;
  int nB = k;
  if (32 + product_bits(q, k) >= 61 * k + 61) ++nB;
  if (nB + 1 > PIRB_MAX_BSK) return false;
  C.nB = nB;
  const std::vector<u64> aux = get_primes(N, 61, (size_t)nB + 2);
  if ((int)aux.size() != nB + 2) return false;
  std::vector<u64> B(aux.begin() + 2, aux.end()), Bsk = B;
  const u64 m_sk = aux[0];
  Bsk.push_back(m_sk);
  for (u64 p : Bsk) {
    if (p == P) return false;
    for (int j = 0; j < k; ++j)
      if (p == q[j]) return false;
  }
  *bsk_out = Bsk;
  const u64 mt = 1ull << 32;
  for (int j = 0; j < k; ++j) {
    C.q[j] = mod_lite(q[j]);
    C.t_mod_q[j] = t % q[j];
    C.mtilde_mod_q[j] = mt % q[j];
    C.inv_qhat_mod_q[j] = invmod_prime(prod_mod(q, k, j, q[j]), q[j]);
    C.qhat_mod_mtilde[j] = prod_mod(q, k, j, mt);
    C.b_mod_q[j] = prod_mod(B.data(), nB, -1, q[j]);
    for (int i = 0; i < nB; ++i) C.bhat_mod_q[j][i] = prod_mod(B.data(), nB, i, q[j]);
    C.half_P_mod_q[j] = (P >> 1) % q[j];
    C.inv_P_mod_q[j] = invmod_prime(P % q[j], q[j]);
  }
  for (int i = 0; i <= nB; ++i) {
    const u64 p = Bsk[i];
    C.bsk[i] = mod_lite(p);
    C.t_mod_bsk[i] = t % p;
    for (int j = 0; j < k; ++j) C.qhat_mod_bsk[i][j] = prod_mod(q, k, j, p);
    C.q_mod_bsk[i] = prod_mod(q, k, -1, p);
    C.inv_mtilde_mod_bsk[i] = invmod_prime(mt % p, p);
    C.inv_q_mod_bsk[i] = invmod_prime(C.q_mod_bsk[i], p);
  }
  {  // -Q^-1 mod 2^32 by Newton iteration (Q is odd)
    const u64 qm = prod_mod(q, k, -1, mt);
    u64 x = qm;
    for (int it = 0; it < 5; ++it) x = (x * (2 - qm * x)) & (mt - 1);
    C.neg_inv_q_mod_mtilde = (mt - x) & (mt - 1);
  }
  for (int j = 0; j < nB; ++j) {
    C.inv_bhat_mod_b[j] = invmod_prime(prod_mod(B.data(), nB, j, B[j]), B[j]);
    C.bhat_mod_msk[j] = prod_mod(B.data(), nB, j, m_sk);
  }
  C.inv_b_mod_msk = invmod_prime(prod_mod(B.data(), nB, -1, m_sk), m_sk);
  C.P = mod_lite(P);
  C.half_P = P >> 1;
  return true;
}

}  // namespace hm
}  // namespace pirb
