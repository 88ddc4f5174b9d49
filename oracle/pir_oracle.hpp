// =============================================================================
// oracle/pir_oracle.hpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A CPU restatement of the OpenMined/PIR server answer path and of the
// Microsoft SEAL 3.5.6 routines that path calls (SEAL is an un-vendored
// third-party dependency of the reference: pinned at pir/deps.bzl:64-71,
// sha256 13674a39..., built by third_party/seal.BUILD:9-29; its source is NOT
// under /root/reference, so the SEAL-internal parts below restate SEAL's
// published algorithms: Harvey-lazy negacyclic NTT with the minimal primitive
// 2N-th root and bit-reversed output, Barrett reduction, RNS special-prime key
// switching with rounding mod-down, BFV encrypt/decrypt).
//
// PARITY STATUS: pinned at the DECRYPT level against every known-answer
// vector of the reference's own tests for this path (tests/test_oracle_kat.py
// lists them with file:line).  Ciphertext-limb parity against SEAL itself is
// "parity unpinned": no reference test fixes ciphertext bytes and SEAL cannot
// be built here (no network, no source).  See DESIGN.md §Oracle.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use this code.  The product (pir_b200/) never
// links, imports or calls it.
// =============================================================================
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

typedef uint64_t u64;
typedef unsigned __int128 u128;

// ----------------------------------------------------------------------------
// Modular arithmetic (SEAL util/uintarithsmallmod.h semantics: canonical
// representatives in [0,q), Barrett with floor(2^128/q)).
// ----------------------------------------------------------------------------
struct Modulus {
  u64 q = 0;
  u64 ratio_lo = 0, ratio_hi = 0;  // floor(2^128 / q)
  int bits = 0;
  Modulus() {}
  explicit Modulus(u64 v) : q(v) {
    // floor(2^128 / q) by long division of (2^128 - 1) then fix-up.
    u128 all = ~(u128)0;  // 2^128 - 1
    u128 r = all / q;
    if ((all % q) == (u128)(q - 1)) r += 1;  // (2^128-1)%q == q-1  <=> q | 2^128 (never for odd q>1)
    ratio_lo = (u64)r;
    ratio_hi = (u64)(r >> 64);
    bits = 64 - __builtin_clzll(q);
  }
};

static inline u64 mulhi(u64 a, u64 b) { return (u64)(((u128)a * b) >> 64); }

// 128-bit -> [0,q).  Restates SEAL barrett_reduce_128.
static inline u64 barrett_reduce_128(u64 lo, u64 hi, const Modulus& m) {
  u64 carry = mulhi(lo, m.ratio_lo);
  u128 t2 = (u128)lo * m.ratio_hi;
  u64 tmp1 = (u64)t2 + carry;
  u64 tmp3 = (u64)(t2 >> 64) + (tmp1 < carry);
  t2 = (u128)hi * m.ratio_lo;
  u64 tmp1b = tmp1 + (u64)t2;
  carry = (u64)(t2 >> 64) + (tmp1b < tmp1);
  u64 quot = hi * m.ratio_hi + tmp3 + carry;
  u64 r = lo - quot * m.q;
  return r >= m.q ? r - m.q : r;
}
// 64-bit -> [0,q).  Restates SEAL barrett_reduce_63 (valid for any 64-bit input here).
static inline u64 barrett_reduce_64(u64 a, const Modulus& m) {
  u64 quot = mulhi(a, m.ratio_hi);
  u64 r = a - quot * m.q;
  while (r >= m.q) r -= m.q;
  return r;
}
static inline u64 mulmod(u64 a, u64 b, const Modulus& m) {
  u128 p = (u128)a * b;
  return barrett_reduce_128((u64)p, (u64)(p >> 64), m);
}
static inline u64 addmod(u64 a, u64 b, const Modulus& m) {
  u64 s = a + b;
  return s >= m.q ? s - m.q : s;
}
static inline u64 submod(u64 a, u64 b, const Modulus& m) { return a >= b ? a - b : a + m.q - b; }
static inline u64 negmod(u64 a, const Modulus& m) { return a ? m.q - a : 0; }
static inline u64 powmod(u64 a, u64 e, const Modulus& m) {
  u64 r = 1;
  a = barrett_reduce_64(a, m);
  while (e) {
    if (e & 1) r = mulmod(r, a, m);
    a = mulmod(a, a, m);
    e >>= 1;
  }
  return r;
}
static inline u64 invmod(u64 a, const Modulus& m) { return powmod(a, m.q - 2, m); }  // q prime

static inline bool is_prime_u64(u64 n) {
  if (n < 2) return false;
  static const u64 small[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
  for (u64 p : small) {
    if (n == p) return true;
    if (n % p == 0) return false;
  }
  u64 d = n - 1;
  int s = 0;
  while (!(d & 1)) { d >>= 1; ++s; }
  Modulus m(n);
  for (u64 a : small) {  // deterministic Miller-Rabin for 64-bit with these bases
    u64 x = powmod(a, d, m);
    if (x == 1 || x == n - 1) continue;
    bool comp = true;
    for (int r = 1; r < s; ++r) {
      x = mulmod(x, x, m);
      if (x == n - 1) { comp = false; break; }
    }
    if (comp) return false;
  }
  return true;
}

static inline uint32_t reverse_bits(uint32_t v, int bits) {
  uint32_t r = 0;
  for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1u) << (bits - 1 - i);
  return r;
}

// ----------------------------------------------------------------------------
// Reference shape math (pir/cpp/utils.h:29-37, utils.cpp:7-44,
// database.cpp:318-342, string_encoder.cpp:25-31, parameters.cpp:56-107).
// ----------------------------------------------------------------------------
template <typename T>
static inline T next_power_two(T n) {  // utils.h:29-37
  if (n == 0) return 1;
  --n;
  for (size_t i = 1; i < sizeof(n) * 8; i <<= 1) n |= n >> i;
  return n + 1;
}
static inline uint32_t ceil_log2_u32(uint32_t v) {  // utils.cpp:30-44 (value semantics)
  // v==0 yields 0 in the reference (wraps through the De Bruijn table).
  if (v <= 1) return 0;
  return 32 - __builtin_clz(v - 1);
}
static inline uint32_t log2_u32(uint32_t v) {  // utils.cpp:16-28: truncated log2 (0 -> 0)
  if (v == 0) return 0;
  return 31 - __builtin_clz(v);
}
static inline std::vector<uint32_t> generate_galois_elts(u64 N) {  // utils.cpp:7-14
  std::vector<uint32_t> e(ceil_log2_u32((uint32_t)N));
  for (size_t i = 0; i < e.size(); ++i) e[i] = (uint32_t)((N >> i) + 1);
  return e;
}
static inline std::vector<uint32_t> calculate_dimensions(uint32_t db_size, uint32_t num_dimensions) {
  // database.cpp:334-342 — identical double-precision operations.
  std::vector<uint32_t> results;
  for (int i = (int)num_dimensions; i > 0; --i) {
    results.push_back((uint32_t)std::ceil(std::pow((double)db_size, 1.0 / i)));
    db_size = (uint32_t)std::ceil(static_cast<double>(db_size) / results.back());
  }
  return results;
}
// PlainModulus::Batching(N, bits) [SEAL get_primes]: largest prime p ≡ 1 mod 2N with
// 2^(bits-1) < p < 2^bits, scanning 2^bits - 2N + 1, then -2N steps.
static inline u64 plain_modulus_batching(u64 N, int bit_size) {
  u64 factor = 2 * N;
  u64 value = ((u64)1 << bit_size) - factor + 1;
  u64 lower = (u64)1 << (bit_size - 1);
  while (value > lower) {
    if (is_prime_u64(value)) return value;
    value -= factor;
  }
  throw std::logic_error("failed to find enough qualifying primes");
}
// CoeffModulus::BFVDefault(N) for 128-bit security [SEAL util/globals.cpp tables].
static inline std::vector<u64> bfv_default_coeff_modulus(u64 N) {
  switch (N) {
    case 1024: return {0x7e00001ULL};
    case 2048: return {0x3fffffff000001ULL};
    case 4096: return {0xffffee001ULL, 0xffffc4001ULL, 0x1ffffe0001ULL};
    case 8192: return {0x7fffffd8001ULL, 0x7fffffc8001ULL, 0xfffffffc001ULL, 0xffffff6c001ULL, 0xfffffebc001ULL};
    case 16384:
      return {0xfffffffd8001ULL, 0xfffffffa0001ULL, 0xfffffff00001ULL, 0x1fffffff68001ULL, 0x1fffffff50001ULL,
              0x1ffffffee8001ULL, 0x1ffffffea0001ULL, 0x1ffffffe88001ULL, 0x1ffffffe48001ULL};
    default: throw std::invalid_argument("no BFVDefault table for this degree");
  }
}

// ----------------------------------------------------------------------------
// NTT tables [SEAL util/ntt.cpp + numth.cpp try_minimal_primitive_root]
// ----------------------------------------------------------------------------
struct NttTable {
  Modulus mod;
  int logn = 0;
  size_t n = 0;
  u64 psi = 0;                   // minimal primitive 2N-th root of unity
  std::vector<u64> rp, rp_s;     // rp[bitrev(i)] = psi^i ; Shoup companion floor(w*2^64/q)
  std::vector<u64> irp, irp_s;   // irp[i] = rp[i]^{-1}
  u64 inv_n = 0, inv_n_s = 0;

  static u64 shoup(u64 w, u64 q) { return (u64)((((u128)w) << 64) / q); }

  void init(u64 q, int logn_) {
    mod = Modulus(q);
    logn = logn_;
    n = (size_t)1 << logn;
    u64 two_n = 2 * n;
    if ((q - 1) % two_n) throw std::invalid_argument("modulus not NTT-friendly");
    // any primitive 2N-th root: x^((q-1)/2N) with order exactly 2N  (<=> its N-th power is -1)
    u64 g = 0;
    for (u64 x = 2;; ++x) {
      u64 c = powmod(x, (q - 1) / two_n, mod);
      if (powmod(c, n, mod) == q - 1) { g = c; break; }
    }
    // minimal one: all primitive 2N-th roots are the odd powers of g
    u64 g2 = mulmod(g, g, mod), cur = g, best = g;
    for (size_t i = 0; i < n; ++i) {
      if (cur < best) best = cur;
      cur = mulmod(cur, g2, mod);
    }
    psi = best;
    rp.assign(n, 0); rp_s.assign(n, 0); irp.assign(n, 0); irp_s.assign(n, 0);
    u64 pw = 1;
    for (size_t i = 0; i < n; ++i) {
      size_t r = reverse_bits((uint32_t)i, logn);
      rp[r] = pw;
      pw = mulmod(pw, psi, mod);
    }
    for (size_t i = 0; i < n; ++i) {
      rp_s[i] = shoup(rp[i], q);
      irp[i] = invmod(rp[i], mod);
      irp_s[i] = shoup(irp[i], q);
    }
    inv_n = invmod((u64)n % q, mod);
    inv_n_s = shoup(inv_n, q);
  }

  // Forward negacyclic NTT, natural order in -> bit-reversed out: out[i] = a(psi^(2*bitrev(i)+1)).
  // Cooley-Tukey with Harvey lazy butterflies ([0,4q) intermediates), canonical output.
  void forward(u64* a) const {
    const u64 q = mod.q, two_q = 2 * q;
    size_t t = n >> 1;
    for (size_t m = 1; m < n; m <<= 1, t >>= 1) {
      for (size_t i = 0; i < m; ++i) {
        const u64 W = rp[m + i], Ws = rp_s[m + i];
        u64* x = a + 2 * i * t;
        u64* y = x + t;
        for (size_t j = 0; j < t; ++j) {
          u64 X = x[j];
          X -= (X >= two_q) ? two_q : 0;
          u64 Y = y[j];
          u64 Q = mulhi(Ws, Y);
          u64 T = Y * W - Q * q;  // [0,2q)
          x[j] = X + T;
          y[j] = X + two_q - T;
        }
      }
    }
    for (size_t i = 0; i < n; ++i) {
      u64 v = a[i];
      v -= (v >= two_q) ? two_q : 0;
      v -= (v >= q) ? q : 0;
      a[i] = v;
    }
  }
  // Inverse (Gentleman-Sande), bit-reversed in -> natural out, scaled by N^{-1}, canonical output.
  void inverse(u64* a) const {
    const u64 q = mod.q, two_q = 2 * q;
    size_t t = 1;
    for (size_t m = n >> 1; m >= 1; m >>= 1, t <<= 1) {
      for (size_t i = 0; i < m; ++i) {
        const u64 W = irp[m + i], Ws = irp_s[m + i];
        u64* x = a + 2 * i * t;
        u64* y = x + t;
        for (size_t j = 0; j < t; ++j) {
          u64 X = x[j], Y = y[j];
          u64 S = X + Y;
          S -= (S >= two_q) ? two_q : 0;
          u64 D = X + two_q - Y;
          u64 Q = mulhi(Ws, D);
          x[j] = S;
          y[j] = D * W - Q * q;
        }
      }
    }
    for (size_t i = 0; i < n; ++i) {
      u64 v = a[i];
      u64 Q = mulhi(inv_n_s, v);
      v = v * inv_n - Q * q;
      v -= (v >= q) ? q : 0;
      a[i] = v;
    }
  }
};

// ----------------------------------------------------------------------------
// Context: N, data moduli q_0..q_{k-1}, special prime P (= last key-level
// modulus), plain modulus t.  Ciphertexts on this path live at the first data
// level (server.cpp:81-85).
// ----------------------------------------------------------------------------
struct Context {
  size_t N = 0;
  int logn = 0;
  size_t k = 0;                 // number of data-level moduli
  std::vector<NttTable> tb;     // k+1 tables; tb[k] is the special prime
  u64 t = 0;
  uint32_t ptb = 0;             // trunc(log2 t)
  std::vector<u64> inv_P_mod_q; // P^{-1} mod q_j
  std::vector<u64> half_P_mod_q;
  u64 half_P = 0;

  Context(size_t N_, const std::vector<u64>& moduli, u64 t_) : N(N_), t(t_) {
    if (moduli.size() < 2) throw std::invalid_argument("need at least one data modulus and a special prime");
    logn = 0;
    while (((size_t)1 << logn) < N) ++logn;
    if (((size_t)1 << logn) != N) throw std::invalid_argument("N must be a power of two");
    k = moduli.size() - 1;
    tb.resize(k + 1);
    for (size_t i = 0; i <= k; ++i) tb[i].init(moduli[i], logn);
    ptb = log2_u32((uint32_t)t);
    u64 P = moduli[k];
    half_P = P >> 1;
    inv_P_mod_q.resize(k);
    half_P_mod_q.resize(k);
    for (size_t j = 0; j < k; ++j) {
      inv_P_mod_q[j] = invmod(barrett_reduce_64(P, tb[j].mod), tb[j].mod);
      half_P_mod_q[j] = barrett_reduce_64(half_P, tb[j].mod);
    }
  }
  const Modulus& mod(size_t j) const { return tb[j].mod; }
  u64 q(size_t j) const { return tb[j].mod.q; }
  u64 P() const { return tb[k].mod.q; }
  size_t ct_limbs() const { return 2 * k * N; }
  size_t pt_limbs() const { return k * N; }

  // ct_reencoder.cpp:29-38 (double log2 / ceil, truncated log2 of t)
  uint32_t expansion_ratio() const {
    uint32_t er = 0;
    for (size_t j = 0; j < k; ++j) er += (uint32_t)std::ceil(std::log2((double)q(j)) / ptb);
    return er;
  }
  uint32_t local_expansion(size_t j) const { return (uint32_t)std::ceil(std::log2((double)q(j)) / ptb); }
};

// ---------------------------------------------------------------------------
// [SEAL GaloisTool::apply_galois, coefficient form]  
// ---------------------------------------------------------------------------
static inline void apply_galois_poly(const u64* in, size_t N, int logn, uint32_t g, const Modulus& m, u64* out) {
  u64 index_raw = 0;
  for (size_t i = 0; i < N; ++i, index_raw += g) {
    size_t idx = index_raw & (N - 1);
    u64 v = in[i];
    if ((index_raw >> logn) & 1) v = negmod(v, m);
    out[idx] = v;
  }
}

// [SEAL util::negacyclic_shift_poly_coeffmod]  
static inline void negacyclic_shift_poly(const u64* in, size_t N, size_t shift, const Modulus& m, u64* out) {
  if (shift == 0) { std::memcpy(out, in, N * sizeof(u64)); return; }
  u64 index_raw = shift;
  for (size_t i = 0; i < N; ++i, ++index_raw) {
    size_t idx = index_raw & (N - 1);
    u64 v = in[i];
    if ((index_raw & N) && v) v = m.q - v;
    out[idx] = v;
  }
}

// server.cpp:78-103 multiply_inverse_power_of_x
static inline void multiply_inverse_power_of_x(const Context& c, const u64* ct_in, uint32_t kpow, u64* ct_out) {
  const size_t N = c.N;
  uint32_t index = (uint32_t)(((N << 1) - kpow) % (N << 1));  // server.cpp:87-88 (uint32 arithmetic)
  for (size_t i = 0; i < 2; ++i)
    for (size_t j = 0; j < c.k; ++j)
      negacyclic_shift_poly(ct_in + (i * c.k + j) * N, N, index, c.mod(j), ct_out + (i * c.k + j) * N);
}

// ---------------------------------------------------------------------------
// Galois keys, raw layout: key for one element = [J<k][c<2][I<k+1][N] limbs, NTT form
// (SEAL KSwitchKeys: keys()[index][J].data().data(c)[I*N + n]).
// ---------------------------------------------------------------------------
struct GaloisKeys {
  std::vector<uint32_t> elts;
  std::vector<u64> data;  // [n_elts][k][2][k+1][N]
  size_t key_limbs(const Context& c) const { return c.k * 2 * (c.k + 1) * c.N; }
  const u64* key_for(const Context& c, uint32_t g) const {
    for (size_t i = 0; i < elts.size(); ++i)
      if (elts[i] == g) return data.data() + i * key_limbs(c);
    return nullptr;
  }
};

// [SEAL Evaluator::switch_key_inplace, BFV branch]  
// target: [k][N] coefficient form.  Adds the switched (c0,c1) contribution into ct ([2][k][N]).
static inline void switch_key_inplace(const Context& c, u64* ct, const u64* target, const u64* key) {
  const size_t N = c.N, k = c.k;
  std::vector<u64> prod(2 * (k + 1) * N);  // [comp][I][N]
  std::vector<u64> tntt(N);
  std::vector<u128> lazy(2 * N);
  for (size_t I = 0; I <= k; ++I) {
    const NttTable& T = c.tb[I];
    std::fill(lazy.begin(), lazy.end(), (u128)0);
    for (size_t J = 0; J < k; ++J) {
      const u64* src = target + J * N;
      if (c.q(J) <= T.mod.q) {
        std::memcpy(tntt.data(), src, N * sizeof(u64));
      } else {
        for (size_t n = 0; n < N; ++n) tntt[n] = barrett_reduce_64(src[n], T.mod);
      }
      T.forward(tntt.data());
      for (size_t comp = 0; comp < 2; ++comp) {
        const u64* kp = key + ((J * 2 + comp) * (k + 1) + I) * N;
        u128* acc = lazy.data() + comp * N;
        for (size_t n = 0; n < N; ++n) acc[n] += (u128)tntt[n] * kp[n];
      }
    }
    for (size_t comp = 0; comp < 2; ++comp) {
      u64* dst = prod.data() + (comp * (k + 1) + I) * N;
      const u128* acc = lazy.data() + comp * N;
      for (size_t n = 0; n < N; ++n) dst[n] = barrett_reduce_128((u64)acc[n], (u64)(acc[n] >> 64), T.mod);
    }
  }
  // mod-down by the special prime with rounding
  const u64 P = c.P();
  for (size_t comp = 0; comp < 2; ++comp) {
    u64* last = prod.data() + (comp * (k + 1) + k) * N;
    c.tb[k].inverse(last);
    for (size_t n = 0; n < N; ++n) {
      u64 v = last[n] + c.half_P;
      last[n] = v >= P ? v - P : v;
    }
    for (size_t j = 0; j < k; ++j) {
      const Modulus& m = c.mod(j);
      u64* pj = prod.data() + (comp * (k + 1) + j) * N;
      c.tb[j].inverse(pj);
      u64* dst = ct + (comp * k + j) * N;
      for (size_t n = 0; n < N; ++n) {
        u64 r = submod(barrett_reduce_64(last[n], m), c.half_P_mod_q[j], m);
        u64 delta = mulmod(submod(pj[n], r, m), c.inv_P_mod_q[j], m);
        dst[n] = addmod(dst[n], delta, m);
      }
    }
  }
}

// [SEAL Evaluator::apply_galois_inplace, BFV]; server.cpp:67-76 substitute_power_x_inplace.
// Returns false if the key for g is missing (SEAL throws -> absl::InternalError).
static inline bool apply_galois_inplace(const Context& c, u64* ct, uint32_t g, const GaloisKeys& gk) {
  const u64* key = gk.key_for(c, g);
  if (!key) return false;
  const size_t N = c.N, k = c.k;
  std::vector<u64> tmp(k * N), target(k * N);
  for (size_t j = 0; j < k; ++j) apply_galois_poly(ct + j * N, N, c.logn, g, c.mod(j), tmp.data() + j * N);
  for (size_t j = 0; j < k; ++j) apply_galois_poly(ct + (k + j) * N, N, c.logn, g, c.mod(j), target.data() + j * N);
  std::memcpy(ct, tmp.data(), k * N * sizeof(u64));
  std::memset(ct + k * N, 0, k * N * sizeof(u64));
  switch_key_inplace(c, ct, target.data(), key);
  return true;
}

static inline void add_ct_inplace(const Context& c, u64* a, const u64* b) {
  for (size_t i = 0; i < 2; ++i)
    for (size_t j = 0; j < c.k; ++j) {
      const Modulus& m = c.mod(j);
      u64* x = a + (i * c.k + j) * c.N;
      const u64* y = b + (i * c.k + j) * c.N;
      for (size_t n = 0; n < c.N; ++n) x[n] = addmod(x[n], y[n], m);
    }
}

// server.cpp:105-146 oblivious_expansion (single ciphertext).
// out: num_items cts ([num_items][2][k][N]).  Returns 0 ok, 3 invalid argument, 13 internal (missing key).
static inline int oblivious_expansion(const Context& c, const u64* ct, size_t num_items, const GaloisKeys& gk,
                                      std::vector<u64>& out) {
  const size_t N = c.N, L = c.ct_limbs();
  if (num_items > N) return 3;  // server.cpp:111-114
  size_t logm = ceil_log2_u32((uint32_t)num_items);
  size_t m = next_power_two(num_items);
  std::vector<u64> res(m * L, 0);
  std::memcpy(res.data(), ct, L * sizeof(u64));
  std::vector<u64> c0(L), c1(L);
  for (size_t j = 0; j < logm; ++j) {
    const size_t two_j = (size_t)1 << j;
    for (size_t kk = 0; kk < two_j; ++kk) {
      u64* rk = res.data() + kk * L;
      u64* rk2 = res.data() + (kk + two_j) * L;
      std::memcpy(c0.data(), rk, L * sizeof(u64));
      if (!apply_galois_inplace(c, c0.data(), (uint32_t)((N >> j) + 1), gk)) return 13;
      multiply_inverse_power_of_x(c, rk, (uint32_t)two_j, rk2);                 // server.cpp:129-130
      multiply_inverse_power_of_x(c, c0.data(), (uint32_t)(N + two_j), c1.data());  // server.cpp:137-138
      add_ct_inplace(c, rk, c0.data());
      add_ct_inplace(c, rk2, c1.data());
    }
  }
  out.assign(res.begin(), res.begin() + num_items * L);  // server.cpp:144
  return 0;
}

// server.cpp:148-171 oblivious_expansion (multi ciphertext)
static inline int oblivious_expansion_multi(const Context& c, const u64* cts, size_t n_ct, size_t total_items,
                                            const GaloisKeys& gk, std::vector<u64>& out) {
  const size_t N = c.N, L = c.ct_limbs();
  if (n_ct != total_items / N + 1) return 3;  // server.cpp:154-158
  out.clear();
  out.reserve(total_items * L);
  for (size_t i = 0; i < n_ct; ++i) {
    std::vector<u64> v;
    int rc = oblivious_expansion(c, cts + i * L, std::min(N, total_items), gk, v);
    if (rc) return rc;
    out.insert(out.end(), v.begin(), v.end());
    total_items -= N;  // wraps on the last iteration exactly like the reference (server.cpp:168)
  }
  return 0;
}

// ---------------------------------------------------------------------------
// Plaintext -> NTT form  [SEAL Evaluator::transform_to_ntt_inplace(Plaintext, parms_id)]
// coeffs: n_coeff <= N values < t.  out: [k][N].
// ---------------------------------------------------------------------------
static inline void plain_to_ntt(const Context& c, const u64* coeffs, size_t n_coeff, u64* out) {
  const u64 thr = (c.t + 1) >> 1;
  for (size_t j = 0; j < c.k; ++j) {
    u64* o = out + j * c.N;
    const u64 inc = c.q(j) - c.t;
    for (size_t n = 0; n < n_coeff; ++n) o[n] = coeffs[n] >= thr ? coeffs[n] + inc : coeffs[n];
    for (size_t n = n_coeff; n < c.N; ++n) o[n] = 0;
    c.tb[j].forward(o);
  }
}
static inline void ct_to_ntt(const Context& c, u64* ct) {
  for (size_t i = 0; i < 2; ++i)
    for (size_t j = 0; j < c.k; ++j) c.tb[j].forward(ct + (i * c.k + j) * c.N);
}
static inline void ct_from_ntt(const Context& c, u64* ct) {
  for (size_t i = 0; i < 2; ++i)
    for (size_t j = 0; j < c.k; ++j) c.tb[j].inverse(ct + (i * c.k + j) * c.N);
}
// multiply_plain, NTT x NTT branch: out = ct ⊙ pt
static inline void multiply_plain_ntt(const Context& c, const u64* ct, const u64* pt, u64* out) {
  for (size_t i = 0; i < 2; ++i)
    for (size_t j = 0; j < c.k; ++j) {
      const Modulus& m = c.mod(j);
      const u64* x = ct + (i * c.k + j) * c.N;
      const u64* y = pt + j * c.N;
      u64* o = out + (i * c.k + j) * c.N;
      for (size_t n = 0; n < c.N; ++n) o[n] = mulmod(x[n], y[n], m);
    }
}

// ct_reencoder.cpp:40-71 Encode: one ct (coefficient form) -> 2*ER plaintexts of N coeffs < 2^ptb
static inline void reencode(const Context& c, const u64* ct, std::vector<u64>& pts /*[2*ER][N]*/) {
  const uint32_t ptb = c.ptb;
  const u64 mask = (u64)(((uint32_t)1 << ptb) - 1);  // 32-bit shift as in ct_reencoder.cpp:45
  const uint32_t ER = c.expansion_ratio();
  pts.assign((size_t)2 * ER * c.N, 0);
  size_t e = 0;
  for (size_t poly = 0; poly < 2; ++poly)
    for (size_t j = 0; j < c.k; ++j) {
      uint32_t le = c.local_expansion(j);
      uint32_t shift = 0;
      for (uint32_t i = 0; i < le; ++i, ++e, shift += ptb) {
        const u64* src = ct + (poly * c.k + j) * c.N;
        u64* dst = pts.data() + e * c.N;
        for (size_t n = 0; n < c.N; ++n) dst[n] = (src[n] >> shift) & mask;
      }
    }
}
// ct_reencoder.cpp:77-112 Decode (client side; harness only)
static inline void reencode_decode(const Context& c, const u64* pts /*[2*ER][N]*/, u64* ct) {
  const uint32_t ptb = c.ptb;
  size_t e = 0;
  for (size_t poly = 0; poly < 2; ++poly)
    for (size_t j = 0; j < c.k; ++j) {
      uint32_t le = c.local_expansion(j);
      uint32_t shift = 0;
      u64* dst = ct + (poly * c.k + j) * c.N;
      for (uint32_t i = 0; i < le; ++i, ++e, shift += ptb) {
        const u64* src = pts + e * c.N;
        for (size_t n = 0; n < c.N; ++n) {
          if (shift == 0) dst[n] = src[n];
          else dst[n] += (src[n] << shift);
        }
      }
    }
}

// ---------------------------------------------------------------------------
// database.cpp:170-258 DatabaseMultiplier::multiply (re-encoder path).
// db: [num_pt][k][N] NTT form.  sv: [dim_sum][2][k][N], coefficient form on entry, transformed
// in place to NTT form when first used (database.cpp:188-191, 221-224) — tracked by sv_ntt flags.
// ---------------------------------------------------------------------------
struct DbMultiplier {
  const Context& c;
  const u64* db;
  size_t num_pt;
  u64* sv;
  std::vector<uint8_t>& sv_ntt;
  size_t db_it = 0;
  uint32_t ER;
  DbMultiplier(const Context& c_, const u64* db_, size_t num_pt_, u64* sv_, std::vector<uint8_t>& flags)
      : c(c_), db(db_), num_pt(num_pt_), sv(sv_), sv_ntt(flags), ER(c_.expansion_ratio()) {}

  // returns result cts (coefficient form), count = (2*ER)^(remaining dims - 1)
  std::vector<u64> multiply(const uint32_t* dims, size_t nd, size_t sv_off) {
    const size_t L = c.ct_limbs();
    const size_t this_dim = dims[0];
    std::vector<u64> result;
    bool first = true;
    std::vector<u64> temp;
    for (size_t i = 0; i < this_dim; ++i) {
      if (db_it == num_pt) break;  // database.cpp:183
      u64* s = sv + (sv_off + i) * L;
      if (nd == 1) {
        temp.resize(L);
        if (!sv_ntt[sv_off + i]) { ct_to_ntt(c, s); sv_ntt[sv_off + i] = 1; }
        multiply_plain_ntt(c, s, db + (db_it++) * c.pt_limbs(), temp.data());
      } else {
        std::vector<u64> lower = multiply(dims + 1, nd - 1, sv_off + this_dim);
        size_t n_lower = lower.size() / L;
        temp.resize(n_lower * 2 * ER * L);
        size_t out_i = 0;
        std::vector<u64> pts, ptntt(c.pt_limbs());
        for (size_t l = 0; l < n_lower; ++l) {
          reencode(c, lower.data() + l * L, pts);
          for (size_t e = 0; e < (size_t)2 * ER; ++e, ++out_i) {
            if (!sv_ntt[sv_off + i]) { ct_to_ntt(c, s); sv_ntt[sv_off + i] = 1; }
            plain_to_ntt(c, pts.data() + e * c.N, c.N, ptntt.data());
            multiply_plain_ntt(c, s, ptntt.data(), temp.data() + out_i * L);
          }
        }
      }
      if (first) {
        result = temp;
        first = false;
      } else {
        for (size_t r = 0; r < result.size() / L; ++r) add_ct_inplace(c, result.data() + r * L, temp.data() + r * L);
      }
    }
    for (size_t r = 0; r < result.size() / L; ++r) ct_from_ntt(c, result.data() + r * L);  // database.cpp:250-254
    return result;
  }
};

// database.cpp:290-316 PIRDatabase::multiply. rc 3 on selection-vector size mismatch.
static inline int db_multiply(const Context& c, const u64* db, size_t num_pt, const uint32_t* dims, size_t nd, u64* sv,
                              size_t n_sv, std::vector<u64>& out) {
  size_t dim_sum = 0;
  for (size_t i = 0; i < nd; ++i) dim_sum += dims[i];
  if (n_sv != dim_sum) return 3;
  std::vector<uint8_t> flags(n_sv, 0);
  DbMultiplier m(c, db, num_pt, sv, flags);
  out = m.multiply(dims, nd, 0);
  return 0;
}

// server.cpp:173-195 processQuery on raw limbs (expand + multiply)
static inline int process_query(const Context& c, const u64* db, size_t num_pt, const uint32_t* dims, size_t nd,
                                const GaloisKeys& gk, const u64* query, size_t n_ct, std::vector<u64>& out) {
  size_t dim_sum = 0;
  for (size_t i = 0; i < nd; ++i) dim_sum += dims[i];
  std::vector<u64> sv;
  int rc = oblivious_expansion_multi(c, query, n_ct, dim_sum, gk, sv);
  if (rc) return rc;
  return db_multiply(c, db, num_pt, dims, nd, sv.data(), dim_sum, out);
}

// ---------------------------------------------------------------------------
// StringEncoder (string_encoder.cpp:58-80, 108-158): MSB-first bit packing.
// ---------------------------------------------------------------------------
static inline size_t string_encode(const uint8_t* bytes, size_t n_bytes, size_t bits_per_coeff, u64* coeffs,
                                   size_t max_coeff) {
  size_t num_coeff = (size_t)std::ceil(static_cast<double>(n_bytes * 8) / bits_per_coeff);  // :86-93
  if (num_coeff > max_coeff) return (size_t)-1;
  for (size_t i = 0; i < num_coeff; ++i) coeffs[i] = 0;
  size_t ci = 0, cb = bits_per_coeff;
  for (size_t b = 0; b < n_bytes; ++b) {
    uint8_t ch = bytes[b];
    size_t remain = 8;
    while (remain > 0) {
      size_t n = std::min(cb, remain);
      coeffs[ci] <<= n;
      coeffs[ci] |= (u64)(ch >> (8 - n));
      ch = (uint8_t)(ch << n);
      cb -= n;
      remain -= n;
      if (cb == 0) { ++ci; cb = bits_per_coeff; }
    }
  }
  if (cb < bits_per_coeff && cb > 0) coeffs[ci] <<= cb;  // terminate()
  return num_coeff;
}
static inline int string_decode(const u64* coeffs, size_t n_coeff, size_t bits_per_coeff, size_t length,
                                size_t byte_offset, uint8_t* out) {
  if ((byte_offset + length) > (n_coeff * bits_per_coeff / 8)) return 3;  // :126-129
  size_t start = byte_offset * 8 / bits_per_coeff;
  size_t cb = ((start + 1) * bits_per_coeff) - (byte_offset * 8);
  if (cb == 0) cb = bits_per_coeff;
  std::memset(out, 0, length);
  size_t ri = 0, remain = 8;
  for (size_t i = start; i < n_coeff; ++i) {
    while (cb > 0) {
      size_t n = std::min(cb, remain);
      out[ri] = (uint8_t)(out[ri] << n);
      out[ri] |= (uint8_t)((coeffs[i] >> (cb - n)) & ((1u << n) - 1));
      cb -= n;
      remain -= n;
      if (remain == 0) {
        if (++ri >= length) return 0;
        remain = 8;
      }
    }
    cb = bits_per_coeff;
  }
  return 0;
}

// ===========================================================================
// HARNESS-ONLY crypto (client side; not on the server path):
// BFV keygen / encrypt / decrypt restated from SEAL 3.5 (ternary secret,
// clipped normal sigma 3.2 / 6 sigma, pk encryption at key level + mod-switch
// to the first data level, Delta*m with rounding correction).
// ===========================================================================
struct Rng {
  u64 s[4];
  explicit Rng(u64 seed) {
    u64 z = seed;
    for (int i = 0; i < 4; ++i) {  // splitmix64
      z += 0x9e3779b97f4a7c15ULL;
      u64 x = z;
      x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
      x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
      s[i] = x ^ (x >> 31);
    }
  }
  static u64 rotl(u64 x, int k) { return (x << k) | (x >> (64 - k)); }
  u64 next() {  // xoshiro256**
    u64 r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
  }
  u64 uniform(u64 bound) {  // rejection sampling
    u64 lim = ~(u64)0 - (~(u64)0 % bound);
    u64 v;
    do v = next(); while (v >= lim);
    return v % bound;
  }
  double unit() { return ((next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
  int64_t noise() {  // clipped normal, sigma 3.2, |x| <= 19.2
    for (;;) {
      double u1 = unit(), u2 = unit();
      double g = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2) * 3.2;
      if (std::fabs(g) <= 19.2) return (int64_t)std::llround(g);
    }
  }
};

struct SecretKey { std::vector<u64> ntt; std::vector<int8_t> coeff; };  // ntt: [k+1][N]
struct PublicKey { std::vector<u64> ntt; };                              // [2][k+1][N] NTT form

static inline void small_to_rns_ntt(const Context& c, const std::vector<int64_t>& v, size_t nmod, u64* out, bool ntt) {
  for (size_t j = 0; j < nmod; ++j) {
    u64* o = out + j * c.N;
    for (size_t n = 0; n < c.N; ++n) o[n] = v[n] >= 0 ? (u64)v[n] : c.q(j) - (u64)(-v[n]);
    if (ntt) c.tb[j].forward(o);
  }
}
static inline SecretKey gen_secret_key(const Context& c, Rng& rng) {
  SecretKey sk;
  sk.coeff.resize(c.N);
  std::vector<int64_t> v(c.N);
  for (size_t n = 0; n < c.N; ++n) { v[n] = (int64_t)rng.uniform(3) - 1; sk.coeff[n] = (int8_t)v[n]; }
  sk.ntt.resize((c.k + 1) * c.N);
  small_to_rns_ntt(c, v, c.k + 1, sk.ntt.data(), true);
  return sk;
}
// (c0,c1) = (-(a*s + e), a) over all k+1 moduli, NTT form [SEAL encrypt_zero_symmetric, is_ntt_form=true]
static inline void encrypt_zero_symmetric_ntt(const Context& c, const SecretKey& sk, Rng& rng, u64* out /*[2][k+1][N]*/) {
  const size_t N = c.N, K1 = c.k + 1;
  std::vector<int64_t> e(N);
  for (size_t n = 0; n < N; ++n) e[n] = rng.noise();
  std::vector<u64> ev(K1 * N);
  small_to_rns_ntt(c, e, K1, ev.data(), true);
  for (size_t j = 0; j < K1; ++j) {
    const Modulus& m = c.mod(j);
    u64* c0 = out + j * N;
    u64* c1 = out + (K1 + j) * N;
    for (size_t n = 0; n < N; ++n) {
      c1[n] = rng.uniform(m.q);
      c0[n] = negmod(addmod(mulmod(c1[n], sk.ntt[j * N + n], m), ev[j * N + n], m), m);
    }
  }
}
static inline PublicKey gen_public_key(const Context& c, const SecretKey& sk, Rng& rng) {
  PublicKey pk;
  pk.ntt.resize(2 * (c.k + 1) * c.N);
  encrypt_zero_symmetric_ntt(c, sk, rng, pk.ntt.data());
  return pk;
}
// [SEAL KeyGenerator::generate_one_kswitch_key + galois_keys]: key for element g.
static inline void gen_galois_key(const Context& c, const SecretKey& sk, uint32_t g, Rng& rng, u64* out /*[k][2][k+1][N]*/) {
  const size_t N = c.N, K1 = c.k + 1;
  // sigma_g(s) in coefficient form, then NTT per key-level modulus
  std::vector<int64_t> rot(N);
  u64 index_raw = 0;
  for (size_t i = 0; i < N; ++i, index_raw += g) {
    size_t idx = index_raw & (N - 1);
    int64_t v = sk.coeff[i];
    if ((index_raw >> c.logn) & 1) v = -v;
    rot[idx] = v;
  }
  std::vector<u64> rot_ntt(K1 * N);
  small_to_rns_ntt(c, rot, K1, rot_ntt.data(), true);
  for (size_t J = 0; J < c.k; ++J) {
    u64* kj = out + J * 2 * K1 * N;
    encrypt_zero_symmetric_ntt(c, sk, rng, kj);
    const Modulus& m = c.mod(J);
    u64 factor = barrett_reduce_64(c.P(), m);
    u64* c0J = kj + J * N;  // comp 0, modulus J
    for (size_t n = 0; n < N; ++n) c0J[n] = addmod(c0J[n], mulmod(rot_ntt[J * N + n], factor, m), m);
  }
}

// Tiny fixed-width unsigned bigint (little-endian 64-bit limbs) for exact decryption rounding.
struct Big {
  static const int W = 12;
  u64 v[W];
  Big() { std::memset(v, 0, sizeof(v)); }
  explicit Big(u64 x) { std::memset(v, 0, sizeof(v)); v[0] = x; }
  void add(const Big& o) { u128 c = 0; for (int i = 0; i < W; ++i) { c += (u128)v[i] + o.v[i]; v[i] = (u64)c; c >>= 64; } }
  void sub(const Big& o) { u64 b = 0; for (int i = 0; i < W; ++i) { u128 d = (u128)v[i] - o.v[i] - b; v[i] = (u64)d; b = (u64)(d >> 64) & 1; } }
  void mul_u64(u64 x) { u128 c = 0; for (int i = 0; i < W; ++i) { c += (u128)v[i] * x; v[i] = (u64)c; c >>= 64; } }
  int cmp(const Big& o) const { for (int i = W - 1; i >= 0; --i) { if (v[i] != o.v[i]) return v[i] < o.v[i] ? -1 : 1; } return 0; }
  u64 mod_u64(u64 m) const { u128 r = 0; for (int i = W - 1; i >= 0; --i) r = ((r << 64) | v[i]) % m; return (u64)r; }
  // floor(this / m) for small m, in place; returns remainder
  u64 div_u64(u64 m) { u128 r = 0; for (int i = W - 1; i >= 0; --i) { u128 cur = (r << 64) | v[i]; v[i] = (u64)(cur / m); r = cur % m; } return (u64)r; }
};

struct Crypto {
  const Context& c;
  std::vector<Big> Qhat;            // Q / q_j
  std::vector<u64> Qhat_inv;        // (Q/q_j)^{-1} mod q_j
  Big Q, twoQ;
  std::vector<u64> delta_mod_q;     // floor(Q/t) mod q_j
  u64 Q_mod_t;
  explicit Crypto(const Context& c_) : c(c_) {
    if (c.k * 64 + 80 > (size_t)Big::W * 64) throw std::invalid_argument("too many moduli for harness bigint");
    Q = Big(1);
    for (size_t j = 0; j < c.k; ++j) Q.mul_u64(c.q(j));
    twoQ = Q; twoQ.add(Q);
    Qhat.resize(c.k); Qhat_inv.resize(c.k);
    for (size_t j = 0; j < c.k; ++j) {
      Big h(1);
      for (size_t i = 0; i < c.k; ++i) if (i != j) h.mul_u64(c.q(i));
      Qhat[j] = h;
      Qhat_inv[j] = invmod(h.mod_u64(c.q(j)), c.mod(j));
    }
    Big d = Q;
    Q_mod_t = d.div_u64(c.t);  // d = floor(Q/t)
    delta_mod_q.resize(c.k);
    for (size_t j = 0; j < c.k; ++j) delta_mod_q[j] = d.mod_u64(c.q(j));
  }

  // Fresh pk encryption of pt (n_coeff coeffs < t) -> ct [2][k][N] coefficient form, first data level.
  void encrypt(const PublicKey& pk, const u64* pt, size_t n_coeff, Rng& rng, u64* ct) const {
    const size_t N = c.N, k = c.k, K1 = k + 1;
    std::vector<int64_t> u(N), e0(N), e1(N);
    for (size_t n = 0; n < N; ++n) { u[n] = (int64_t)rng.uniform(3) - 1; e0[n] = rng.noise(); e1[n] = rng.noise(); }
    std::vector<u64> un(K1 * N), tmp(2 * K1 * N);
    small_to_rns_ntt(c, u, K1, un.data(), true);
    for (size_t comp = 0; comp < 2; ++comp) {
      const std::vector<int64_t>& e = comp ? e1 : e0;
      for (size_t j = 0; j < K1; ++j) {
        const Modulus& m = c.mod(j);
        u64* o = tmp.data() + (comp * K1 + j) * N;
        const u64* p = pk.ntt.data() + (comp * K1 + j) * N;
        for (size_t n = 0; n < N; ++n) o[n] = mulmod(p[n], un[j * N + n], m);
        c.tb[j].inverse(o);
        for (size_t n = 0; n < N; ++n) {
          u64 ev = e[n] >= 0 ? (u64)e[n] : m.q - (u64)(-e[n]);
          o[n] = addmod(o[n], ev, m);
        }
      }
      // mod-switch key level -> first data level [SEAL divide_and_round_q_last_inplace]
      const u64 P = c.P();
      u64* last = tmp.data() + (comp * K1 + k) * N;
      for (size_t n = 0; n < N; ++n) { u64 v = last[n] + c.half_P; last[n] = v >= P ? v - P : v; }
      for (size_t j = 0; j < k; ++j) {
        const Modulus& m = c.mod(j);
        const u64* src = tmp.data() + (comp * K1 + j) * N;
        u64* dst = ct + (comp * k + j) * N;
        for (size_t n = 0; n < N; ++n) {
          u64 r = submod(barrett_reduce_64(last[n], m), c.half_P_mod_q[j], m);
          dst[n] = mulmod(submod(src[n], r, m), c.inv_P_mod_q[j], m);
        }
      }
    }
    // c0 += round(Q*m/t)  [SEAL multiply_add_plain_with_scaling_variant]
    const u64 thr = (c.t + 1) >> 1;
    for (size_t n = 0; n < n_coeff; ++n) {
      u128 num = (u128)pt[n] * Q_mod_t + thr;
      u64 fix = (u64)(num / c.t);
      for (size_t j = 0; j < k; ++j) {
        const Modulus& m = c.mod(j);
        u64 scaled = addmod(mulmod(barrett_reduce_64(pt[n], m), delta_mod_q[j], m), barrett_reduce_64(fix, m), m);
        ct[j * N + n] = addmod(ct[j * N + n], scaled, m);
      }
    }
  }

  // ct [2][k][N] coefficient form -> pt[N] = round(t*(c0 + c1*s)/Q) mod t ; also returns noise budget bits.
  int decrypt(const SecretKey& sk, const u64* ct, u64* pt) const {
    const size_t N = c.N, k = c.k;
    std::vector<u64> phase(k * N);
    for (size_t j = 0; j < k; ++j) {
      const Modulus& m = c.mod(j);
      u64* p = phase.data() + j * N;
      std::memcpy(p, ct + (k + j) * N, N * sizeof(u64));
      c.tb[j].forward(p);
      for (size_t n = 0; n < N; ++n) p[n] = mulmod(p[n], sk.ntt[j * N + n], m);
      c.tb[j].inverse(p);
      for (size_t n = 0; n < N; ++n) p[n] = addmod(p[n], ct[j * N + n], m);
    }
    // invariant noise: v = t*x mod Q centred; budget = log2(Q) - log2(|v|) - 1
    Big max_noise;
    Big halfQ = Q; halfQ.div_u64(2);
    for (size_t n = 0; n < N; ++n) {
      // x = sum_j y_j * Qhat_j mod Q
      Big x;
      u64 acc = 0;  // running integer part of sum_j t*y_j/q_j, mod t
      Big frac;     // sum_j r_j * Qhat_j
      for (size_t j = 0; j < k; ++j) {
        const Modulus& m = c.mod(j);
        u64 y = mulmod(phase[j * N + n], Qhat_inv[j], m);
        u128 ty = (u128)y * c.t;
        u64 a = (u64)(ty / m.q), r = (u64)(ty % m.q);
        acc = (u64)(((u128)acc + a) % c.t);
        Big term = Qhat[j];
        term.mul_u64(r);
        frac.add(term);
      }
      // frac/Q = fractional sum in [0,k); result = acc + floor(frac/Q + 1/2); noise = t*x mod Q = frac mod Q (centred)
      Big w = frac; w.add(frac); w.add(Q);  // 2*frac + Q
      u64 carry = 0;
      while (w.cmp(twoQ) >= 0) { w.sub(twoQ); ++carry; }
      pt[n] = (u64)(((u128)acc + carry) % c.t);
      Big v = frac;
      while (v.cmp(Q) >= 0) v.sub(Q);
      if (v.cmp(halfQ) > 0) { Big z = Q; z.sub(v); v = z; }
      if (v.cmp(max_noise) > 0) max_noise = v;
    }
    auto bitlen = [](const Big& b) { for (int i = Big::W - 1; i >= 0; --i) if (b.v[i]) return i * 64 + 64 - __builtin_clzll(b.v[i]); return 0; };
    int budget = bitlen(Q) - bitlen(max_noise) - 1;
    return budget < 0 ? 0 : budget;
  }
};

}  // namespace orc

#include "bfv_mul_oracle.hpp"  // ciphertext-multiplication mode (database.cpp:202-211)
