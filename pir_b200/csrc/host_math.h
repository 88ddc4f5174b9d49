// host_math.h — host-side number theory for building the device tables (product code; independent of oracle/).
// Restates what SEAL 3.5.6 computes when a SEALContext is created: minimal primitive 2N-th root
// (util/numth.cpp try_minimal_primitive_root), bit-reversed power tables (util/ntt.cpp), Barrett ratios.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace pirb {
namespace hm {

typedef unsigned long long u64;
typedef unsigned __int128 u128;

inline u64 mulmod(u64 a, u64 b, u64 q) { return (u64)((u128)a * b % q); }
inline u64 powmod(u64 a, u64 e, u64 q) {
  u64 r = 1 % q;
  a %= q;
  while (e) {
    if (e & 1) r = mulmod(r, a, q);
    a = mulmod(a, a, q);
    e >>= 1;
  }
  return r;
}
inline u64 invmod_prime(u64 a, u64 q) { return powmod(a, q - 2, q); }
inline u64 shoup(u64 w, u64 q) { return (u64)((((u128)w) << 64) / q); }

inline bool is_prime(u64 n) {
  if (n < 2) return false;
  static const u64 bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
  for (u64 p : bases) {
    if (n == p) return true;
    if (n % p == 0) return false;
  }
  u64 d = n - 1;
  int s = 0;
  while (!(d & 1)) { d >>= 1; ++s; }
  for (u64 a : bases) {
    u64 x = powmod(a, d, n);
    if (x == 1 || x == n - 1) continue;
    bool composite = true;
    for (int r = 1; r < s; ++r) {
      x = mulmod(x, x, n);
      if (x == n - 1) { composite = false; break; }
    }
    if (composite) return false;
  }
  return true;
}

inline uint32_t bitrev(uint32_t v, int bits) {
  uint32_t r = 0;
  for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1u) << (bits - 1 - i);
  return r;
}

// floor(2^128 / q) as (hi, lo)
inline void barrett_ratio(u64 q, u64* hi, u64* lo) {
  u128 all = ~(u128)0;
  u128 r = all / q;  // q odd > 1 never divides 2^128, so floor((2^128-1)/q) == floor(2^128/q)
  *hi = (u64)(r >> 64);
  *lo = (u64)r;
}

// minimal primitive 2N-th root of unity mod prime q (q ≡ 1 mod 2N)
inline u64 minimal_primitive_root(u64 q, u64 N) {
  const u64 two_n = 2 * N;
  u64 g = 0;
  for (u64 x = 2;; ++x) {
    u64 c = powmod(x, (q - 1) / two_n, q);
    if (powmod(c, N, q) == q - 1) { g = c; break; }
  }
  const u64 g2 = mulmod(g, g, q);
  u64 cur = g, best = g;
  for (u64 i = 0; i < N; ++i) {
    if (cur < best) best = cur;
    cur = mulmod(cur, g2, q);
  }
  return best;
}

struct Tables {
  std::vector<u64> rp, rps, irp, irps;
  u64 inv_n, inv_n_s, psi;
  // FP64 engine tables (exact integers as doubles and their x/q companions)
  std::vector<double> fw, fwi, iw, iwi, fin, fini;
};
inline Tables build_tables(u64 q, int logn) {
  Tables T;
  const u64 N = 1ull << logn;
  T.psi = minimal_primitive_root(q, N);
  T.rp.assign(N, 0); T.rps.assign(N, 0); T.irp.assign(N, 0); T.irps.assign(N, 0);
  const u64 psi_inv = invmod_prime(T.psi, q);
  u64 pw = 1, ipw = 1;
  for (u64 i = 0; i < N; ++i) {
    const uint32_t r = bitrev((uint32_t)i, logn);
    T.rp[r] = pw;
    T.irp[r] = ipw;  // (psi^i)^{-1}
    pw = mulmod(pw, T.psi, q);
    ipw = mulmod(ipw, psi_inv, q);
  }
  for (u64 i = 0; i < N; ++i) {
    T.rps[i] = shoup(T.rp[i], q);
    T.irps[i] = shoup(T.irp[i], q);
  }
  T.inv_n = invmod_prime(N % q, q);
  T.inv_n_s = shoup(T.inv_n, q);
  if (q >> 52) return T;  // no FP64 tables for moduli that are not exactly representable
  const double qd = (double)q;
  T.fw.assign(N, 0); T.fwi.assign(N, 0); T.iw.assign(N, 0); T.iwi.assign(N, 0); T.fin.assign(N, 0); T.fini.assign(N, 0);
  for (u64 i = 0; i < N; ++i) {
    T.fw[i] = (double)T.rp[i];
    T.fwi[i] = (double)T.rp[i] / qd;
  }
  // iw[g + j] = psi^(-j*N/g): cyclic inverse DFT twiddles (root psi^-2) of the stage with gap g
  for (u64 g = 1; g < N; g <<= 1) {
    const u64 step = powmod(psi_inv, N / g, q);
    u64 cur = 1;
    for (u64 j = 0; j < g; ++j) {
      T.iw[g + j] = (double)cur;
      T.iwi[g + j] = (double)cur / qd;
      cur = mulmod(cur, step, q);
    }
  }
  u64 cur = T.inv_n;  // fin[i] = N^-1 * psi^-i
  for (u64 i = 0; i < N; ++i) {
    T.fin[i] = (double)cur;
    T.fini[i] = (double)cur / qd;
    cur = mulmod(cur, psi_inv, q);
  }
  return T;
}

// ---- reference shape math (pir/cpp/utils.h:29-37, utils.cpp:30-44, database.cpp:334-342) -------------------
inline u64 next_power_two(u64 n) {
  if (n == 0) return 1;
  --n;
  for (unsigned i = 1; i < 64; i <<= 1) n |= n >> i;
  return n + 1;
}
inline uint32_t ceil_log2(uint32_t v) { return v <= 1 ? 0 : 32 - __builtin_clz(v - 1); }
inline uint32_t trunc_log2(uint32_t v) { return v == 0 ? 0 : 31 - __builtin_clz(v); }
inline std::vector<uint32_t> calculate_dimensions(uint32_t db_size, uint32_t nd) {
  std::vector<uint32_t> r;
  for (int i = (int)nd; i > 0; --i) {
    r.push_back((uint32_t)std::ceil(std::pow((double)db_size, 1.0 / i)));
    db_size = (uint32_t)std::ceil(static_cast<double>(db_size) / r.back());
  }
  return r;
}

}  // namespace hm
}  // namespace pirb
