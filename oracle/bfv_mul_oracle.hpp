// =============================================================================
// oracle/bfv_mul_oracle.hpp — CPU ORACLE, continued.  TEST INFRASTRUCTURE ONLY (see pir_oracle.hpp).
//
// The ciphertext-multiplication mode of the reference (PIRParameters.use_ciphertext_multiplication,
// database.cpp:202-211): upper dimensions multiply the lower result with the selection ciphertext by
// Evaluator::multiply and, when the request carries relinearization keys, Evaluator::relinearize_inplace
// (server.cpp:185-190).  Both live in Microsoft SEAL 3.5.6 (not under /root/reference); restated here from its
// published algorithm, the BEHZ RNS variant of BFV multiplication:
//   util/numth.cpp get_primes          auxiliary 61-bit primes 2^61 - i*2N + 1, descending
//   util/rns.cpp   RNSTool::initialize  bases q, B, Bsk = B u {m_sk}, m_tilde = 2^32 (gamma only serves decryption)
//                  BaseConverter::fast_convert_array, fastbconv_m_tilde, sm_mrq, fast_floor, fastbconv_sk
//   evaluator.cpp  bfv_multiply steps (1)-(8), relinearize_internal -> switch_key_inplace
// Every step is an exact function of canonical residues, so two correct implementations agree limb for limb; which
// residues the *approximate* base conversions produce depends on the auxiliary primes, hence those follow SEAL.
// PARITY STATUS: decrypt level (the reference's CTMultiply cases, tests/test_oracle_kat.py); limb level vs SEAL unpinned.
// =============================================================================
#pragma once

namespace orc {

// [SEAL util/numth.cpp get_primes(ntt_size, bit_size, count)]
static inline std::vector<u64> get_primes(u64 N, int bit_size, size_t count) {
  std::vector<u64> out;
  const u64 factor = 2 * N;
  u64 value = ((u64)1 << bit_size) - factor + 1;
  const u64 lower = (u64)1 << (bit_size - 1);
  while (count > 0 && value > lower) {
    if (is_prime_u64(value)) { out.push_back(value); --count; }
    value -= factor;
  }
  if (count) throw std::logic_error("failed to find enough qualifying primes");
  return out;
}

static inline int product_bit_count(const std::vector<u64>& primes) {
  std::vector<u64> v{1};
  for (u64 p : primes) {
    u128 carry = 0;
    for (auto& limb : v) { carry += (u128)limb * p; limb = (u64)carry; carry >>= 64; }
    if (carry) v.push_back((u64)carry);
  }
  while (v.size() > 1 && !v.back()) v.pop_back();
  return (int)(v.size() - 1) * 64 + (64 - __builtin_clzll(v.back()));
}

// [SEAL util/rns.cpp RNSTool] for the first data level (the level every ciphertext of this path lives at).
struct RnsTool {
  size_t N = 0, k = 0, nB = 0;   // base q size, base B size; Bsk has nB + 1 primes, m_sk last
  std::vector<NttTable> bsk;     // NTT tables of Bsk
  u64 m_sk = 0, gamma = 0;
  static constexpr u64 m_tilde = (u64)1 << 32;
  std::vector<u64> mtilde_mod_q;                 // [j]   m_tilde mod q_j
  std::vector<u64> inv_qhat_mod_q;               // [j]   (Q/q_j)^-1 mod q_j
  std::vector<std::vector<u64>> qhat_mod_bsk;    // [i][j] (Q/q_j) mod bsk_i
  std::vector<u64> qhat_mod_mtilde;              // [j]
  u64 neg_inv_q_mod_mtilde = 0;                  // -Q^-1 mod m_tilde
  std::vector<u64> q_mod_bsk, inv_mtilde_mod_bsk, inv_q_mod_bsk;  // [i]
  std::vector<u64> inv_bhat_mod_b;               // [j]   (B/b_j)^-1 mod b_j
  std::vector<std::vector<u64>> bhat_mod_q;      // [i][j] (B/b_j) mod q_i
  std::vector<u64> bhat_mod_msk;                 // [j]
  u64 inv_b_mod_msk = 0;                         // B^-1 mod m_sk
  std::vector<u64> b_mod_q;                      // [i]   B mod q_i

  static u64 prod_mod(const std::vector<u64>& primes, size_t skip, u64 p) {
    u64 r = 1 % p;
    for (size_t i = 0; i < primes.size(); ++i)
      if (i != skip) r = (u64)((u128)r * (primes[i] % p) % p);
    return r;
  }
  static u64 inv_prime(u64 a, u64 p) { return invmod(a % p, Modulus(p)); }

  explicit RnsTool(const Context& c) : N(c.N), k(c.k) {
    std::vector<u64> q(k);
    for (size_t j = 0; j < k; ++j) q[j] = c.q(j);
    // RNSTool::initialize: |B| = |q|, one more if m_tilde * Q does not fit 61 * |q| + 61 bits
    nB = k;
    if (32 + product_bit_count(q) >= 61 * (int)k + 61) ++nB;
    const std::vector<u64> aux = get_primes(N, 61, nB + 2);  // m_sk, gamma, then B
    m_sk = aux[0];
    gamma = aux[1];
    std::vector<u64> B(aux.begin() + 2, aux.end());
    std::vector<u64> Bsk = B;
    Bsk.push_back(m_sk);
    bsk.resize(nB + 1);
    for (size_t i = 0; i <= nB; ++i) bsk[i].init(Bsk[i], c.logn);

    mtilde_mod_q.resize(k);
    inv_qhat_mod_q.resize(k);
    qhat_mod_mtilde.resize(k);
    for (size_t j = 0; j < k; ++j) {
      mtilde_mod_q[j] = m_tilde % q[j];
      inv_qhat_mod_q[j] = inv_prime(prod_mod(q, j, q[j]), q[j]);
      qhat_mod_mtilde[j] = prod_mod(q, j, m_tilde);
    }
    qhat_mod_bsk.assign(nB + 1, std::vector<u64>(k));
    q_mod_bsk.resize(nB + 1);
    inv_mtilde_mod_bsk.resize(nB + 1);
    inv_q_mod_bsk.resize(nB + 1);
    for (size_t i = 0; i <= nB; ++i) {
      for (size_t j = 0; j < k; ++j) qhat_mod_bsk[i][j] = prod_mod(q, j, Bsk[i]);
      q_mod_bsk[i] = prod_mod(q, (size_t)-1, Bsk[i]);
      inv_mtilde_mod_bsk[i] = inv_prime(m_tilde, Bsk[i]);
      inv_q_mod_bsk[i] = inv_prime(q_mod_bsk[i], Bsk[i]);
    }
    {  // -Q^-1 mod 2^32 (Q odd): Newton iteration
      const u64 qm = prod_mod(q, (size_t)-1, m_tilde);
      u64 x = qm;
      for (int it = 0; it < 5; ++it) x = (x * (2 - qm * x)) & (m_tilde - 1);
      neg_inv_q_mod_mtilde = (m_tilde - x) & (m_tilde - 1);
    }
    inv_bhat_mod_b.resize(nB);
    bhat_mod_msk.resize(nB);
    for (size_t j = 0; j < nB; ++j) {
      inv_bhat_mod_b[j] = inv_prime(prod_mod(B, j, B[j]), B[j]);
      bhat_mod_msk[j] = prod_mod(B, j, m_sk);
    }
    bhat_mod_q.assign(k, std::vector<u64>(nB));
    b_mod_q.resize(k);
    for (size_t i = 0; i < k; ++i) {
      for (size_t j = 0; j < nB; ++j) bhat_mod_q[i][j] = prod_mod(B, j, q[i]);
      b_mod_q[i] = prod_mod(B, (size_t)-1, q[i]);
    }
    inv_b_mod_msk = inv_prime(prod_mod(B, (size_t)-1, m_sk), m_sk);
  }
  size_t n_bsk() const { return nB + 1; }
};

// sum_j v[j] * w[j] mod p for residues below 2^61 [SEAL dot_product_mod: lazy 128-bit sum, one reduction]
static inline u64 dot_mod(const u64* v, const u64* w, size_t n, const Modulus& p) {
  u128 acc = 0;
  for (size_t j = 0; j < n; ++j) {
    acc += (u128)v[j] * w[j];
    if ((j & 15) == 15) acc = barrett_reduce_128((u64)acc, (u64)(acc >> 64), p);
  }
  return barrett_reduce_128((u64)acc, (u64)(acc >> 64), p);
}

// bfv_multiply steps (1)-(2): one polynomial x [k][N] in base q (coefficient form) -> base Bsk [nB+1][N]
// (1) fastbconv_m_tilde: x * m_tilde, fast base conversion q -> Bsk u {m_tilde}
// (2) sm_mrq: Montgomery reduction modulo m_tilde removes the q-overflow of the fast conversion
static inline void behz_extend(const Context& c, const RnsTool& R, const u64* x, u64* y) {
  const size_t N = c.N, k = c.k, nb = R.n_bsk();
  std::vector<u64> tmp(k);
  const u64 mt_mask = RnsTool::m_tilde - 1, mt_half = RnsTool::m_tilde >> 1;
  for (size_t n = 0; n < N; ++n) {
    for (size_t j = 0; j < k; ++j) {
      const Modulus& m = c.mod(j);
      tmp[j] = mulmod(mulmod(x[j * N + n], R.mtilde_mod_q[j], m), R.inv_qhat_mod_q[j], m);
    }
    u64 r_mt = 0;  // conversion to m_tilde = 2^32: wrap-around arithmetic is the reduction
    for (size_t j = 0; j < k; ++j) r_mt += tmp[j] * R.qhat_mod_mtilde[j];
    r_mt = ((r_mt & mt_mask) * R.neg_inv_q_mod_mtilde) & mt_mask;
    for (size_t i = 0; i < nb; ++i) {
      const Modulus& p = R.bsk[i].mod;
      const u64 conv = dot_mod(tmp.data(), R.qhat_mod_bsk[i].data(), k, p);
      u64 r = r_mt;
      if (r >= mt_half) r += p.q - RnsTool::m_tilde;  // centred representative of r modulo m_tilde
      const u128 s = (u128)r * R.q_mod_bsk[i] + conv;
      y[i * N + n] = mulmod(barrett_reduce_128((u64)s, (u64)(s >> 64), p), R.inv_mtilde_mod_bsk[i], p);
    }
  }
}

// bfv_multiply steps (6)-(8) for one output polynomial: dq [k][N], db [nB+1][N] (coefficient form) -> out [k][N]
static inline void behz_scale_and_round(const Context& c, const RnsTool& R, const u64* dq, const u64* db, u64* out) {
  const size_t N = c.N, k = c.k, nB = R.nB, nb = R.n_bsk();
  std::vector<u64> u(k), f(nb), v(nB);
  const Modulus msk = R.bsk[nB].mod;
  const u64 msk_half = R.m_sk >> 1;
  for (size_t n = 0; n < N; ++n) {
    // (6) multiply by t; (7) fast_floor: (t*d - FastBConv_{q->Bsk}(t*d mod q)) / Q  in base Bsk
    for (size_t j = 0; j < k; ++j) {
      const Modulus& m = c.mod(j);
      u[j] = mulmod(mulmod(dq[j * N + n], c.t, m), R.inv_qhat_mod_q[j], m);
    }
    for (size_t i = 0; i < nb; ++i) {
      const Modulus& p = R.bsk[i].mod;
      const u64 conv = dot_mod(u.data(), R.qhat_mod_bsk[i].data(), k, p);
      const u64 tb = mulmod(db[i * N + n], c.t, p);
      f[i] = mulmod(tb + (p.q - conv), R.inv_q_mod_bsk[i], p);
    }
    // (8) fastbconv_sk: B -> q with the Shenoy-Kumaresan correction read off the m_sk residue
    for (size_t j = 0; j < nB; ++j) v[j] = mulmod(f[j], R.inv_bhat_mod_b[j], R.bsk[j].mod);
    const u64 a_sk = dot_mod(v.data(), R.bhat_mod_msk.data(), nB, msk);
    const u64 alpha = mulmod(a_sk + (R.m_sk - f[nB]), R.inv_b_mod_msk, msk);
    for (size_t i = 0; i < k; ++i) {
      const Modulus& m = c.mod(i);
      const u64 conv = dot_mod(v.data(), R.bhat_mod_q[i].data(), nB, m);
      u128 s;
      if (alpha > msk_half) s = (u128)(R.m_sk - alpha) * R.b_mod_q[i] + conv;  // alpha stands for a negative value
      else s = (u128)alpha * (m.q - R.b_mod_q[i]) + conv;
      out[i * N + n] = barrett_reduce_128((u64)s, (u64)(s >> 64), m);
    }
  }
}

// [SEAL Evaluator::bfv_multiply] a: s1 polynomials, b: s2 polynomials, both [.][k][N] coefficient form at the first
// data level.  Returns s1 + s2 - 1 polynomials.
static inline std::vector<u64> bfv_multiply(const Context& c, const RnsTool& R, const u64* a, size_t s1, const u64* b,
                                            size_t s2) {
  const size_t N = c.N, k = c.k, nb = R.n_bsk(), sd = s1 + s2 - 1;
  auto lift = [&](const u64* x, size_t s, std::vector<u64>& xq, std::vector<u64>& xb) {
    xq.assign(x, x + s * k * N);
    xb.assign(s * nb * N, 0);
    for (size_t p = 0; p < s; ++p) {
      behz_extend(c, R, x + p * k * N, xb.data() + p * nb * N);
      for (size_t j = 0; j < k; ++j) c.tb[j].forward(xq.data() + (p * k + j) * N);
      for (size_t i = 0; i < nb; ++i) R.bsk[i].forward(xb.data() + (p * nb + i) * N);
    }
  };
  std::vector<u64> aq, ab, bq, bb;
  lift(a, s1, aq, ab);
  lift(b, s2, bq, bb);
  // (4) D_i = sum_{x+y=i} A_x * B_y in both bases, (5) back to coefficient form
  std::vector<u64> dq(sd * k * N, 0), db(sd * nb * N, 0);
  for (size_t x = 0; x < s1; ++x)
    for (size_t y = 0; y < s2; ++y) {
      for (size_t j = 0; j < k; ++j) {
        const Modulus& m = c.mod(j);
        u64* d = dq.data() + ((x + y) * k + j) * N;
        const u64 *pa = aq.data() + (x * k + j) * N, *pb = bq.data() + (y * k + j) * N;
        for (size_t n = 0; n < N; ++n) d[n] = addmod(d[n], mulmod(pa[n], pb[n], m), m);
      }
      for (size_t i = 0; i < nb; ++i) {
        const Modulus& m = R.bsk[i].mod;
        u64* d = db.data() + ((x + y) * nb + i) * N;
        const u64 *pa = ab.data() + (x * nb + i) * N, *pb = bb.data() + (y * nb + i) * N;
        for (size_t n = 0; n < N; ++n) d[n] = addmod(d[n], mulmod(pa[n], pb[n], m), m);
      }
    }
  std::vector<u64> out(sd * k * N);
  for (size_t p = 0; p < sd; ++p) {
    for (size_t j = 0; j < k; ++j) c.tb[j].inverse(dq.data() + (p * k + j) * N);
    for (size_t i = 0; i < nb; ++i) R.bsk[i].inverse(db.data() + (p * nb + i) * N);
    behz_scale_and_round(c, R, dq.data() + p * k * N, db.data() + p * nb * N, out.data() + p * k * N);
  }
  return out;
}

// [SEAL Evaluator::relinearize_internal, destination size 2] with the one key KeyGenerator::relin_keys() makes:
// a size-3 ciphertext loses its third polynomial through switch_key_inplace.  ct: [3][k][N] -> first [2][k][N] updated.
static inline void relinearize_inplace(const Context& c, u64* ct, const u64* relin_key) {
  switch_key_inplace(c, ct, ct + 2 * c.k * c.N, relin_key);
}

// [SEAL KeyGenerator::relin_keys -> generate_kswitch_keys(s^2)]: out [k][2][k+1][N], same layout as one Galois key
static inline void gen_relin_key(const Context& c, const SecretKey& sk, Rng& rng, u64* out) {
  const size_t N = c.N, K1 = c.k + 1;
  std::vector<u64> s2(K1 * N);
  for (size_t j = 0; j < K1; ++j)
    for (size_t n = 0; n < N; ++n) s2[j * N + n] = mulmod(sk.ntt[j * N + n], sk.ntt[j * N + n], c.mod(j));
  for (size_t J = 0; J < c.k; ++J) {
    u64* kj = out + J * 2 * K1 * N;
    encrypt_zero_symmetric_ntt(c, sk, rng, kj);
    const Modulus& m = c.mod(J);
    const u64 factor = barrett_reduce_64(c.P(), m);
    u64* c0J = kj + J * N;
    for (size_t n = 0; n < N; ++n) c0J[n] = addmod(c0J[n], mulmod(s2[J * N + n], factor, m), m);
  }
}

// Decryption of a ciphertext of `polys` >= 2 polynomials [SEAL Decryptor::bfv_decrypt / dot_product_ct_sk_array]:
// c0 + c1 s + c2 s^2 + ... folded into an equivalent two-polynomial phase input for Crypto::decrypt.
static inline int decrypt_any(const Crypto& crypto, const SecretKey& sk, const u64* ct, size_t polys, u64* pt) {
  const Context& c = crypto.c;
  const size_t N = c.N, k = c.k;
  if (polys == 2) return crypto.decrypt(sk, ct, pt);
  // Horner over the upper polynomials in NTT form: h = c1 + c2 s + ... ; then (c0, h) decrypts as a size-2 object
  std::vector<u64> folded(2 * k * N), acc(N), cur(N);
  std::memcpy(folded.data(), ct, k * N * sizeof(u64));
  for (size_t j = 0; j < k; ++j) {
    const Modulus& m = c.mod(j);
    std::memcpy(acc.data(), ct + ((polys - 1) * k + j) * N, N * sizeof(u64));
    c.tb[j].forward(acc.data());
    for (size_t p = polys - 1; p-- > 1;) {
      std::memcpy(cur.data(), ct + (p * k + j) * N, N * sizeof(u64));
      c.tb[j].forward(cur.data());
      for (size_t n = 0; n < N; ++n) acc[n] = addmod(mulmod(acc[n], sk.ntt[j * N + n], m), cur[n], m);
    }
    c.tb[j].inverse(acc.data());
    std::memcpy(folded.data() + (k + j) * N, acc.data(), N * sizeof(u64));
  }
  return crypto.decrypt(sk, folded.data(), pt);
}

// database.cpp:170-258 with ct_reencoder_ == nullptr (ciphertext-multiplication mode).
// db: [num_pt][k][N] NTT form — the reference keeps the plaintexts of this mode in coefficient form
// (database.cpp:73-76, 102-106) and multiply_plain transforms both operands per call [SEAL multiply_plain_normal]; the
// product is the same canonical polynomial either way.  The selection vector stays in coefficient form (database.cpp:188).
struct CtDbMultiplier {
  const Context& c;
  const RnsTool& R;
  const u64* db;
  size_t num_pt;
  const u64* sv;
  const u64* relin_key;  // nullptr: no relinearization (server.cpp:185-190)
  size_t db_it = 0;
  int error = 0;
  CtDbMultiplier(const Context& c_, const RnsTool& R_, const u64* db_, size_t num_pt_, const u64* sv_, const u64* relin)
      : c(c_), R(R_), db(db_), num_pt(num_pt_), sv(sv_), relin_key(relin) {}

  // returns ONE ciphertext of *polys polynomials (empty if the database ended before this call)
  std::vector<u64> multiply(const uint32_t* dims, size_t nd, size_t sv_off, size_t* polys) {
    const size_t N = c.N, k = c.k, L = c.ct_limbs();
    const size_t this_dim = dims[0];
    std::vector<u64> result, temp;
    size_t rp = 0;
    for (size_t i = 0; i < this_dim; ++i) {
      if (db_it == num_pt) break;  // database.cpp:183
      const u64* s = sv + (sv_off + i) * L;
      size_t tp;
      if (nd == 1) {
        temp.assign(s, s + L);
        ct_to_ntt(c, temp.data());
        multiply_plain_ntt(c, temp.data(), db + (db_it++) * c.pt_limbs(), temp.data());
        ct_from_ntt(c, temp.data());
        tp = 2;
      } else {
        size_t lp = 0;
        std::vector<u64> lower = multiply(dims + 1, nd - 1, sv_off + this_dim, &lp);
        if (error) return {};
        temp = bfv_multiply(c, R, lower.data(), lp, s, 2);  // database.cpp:204-205
        tp = lp + 1;
        if (relin_key) {  // database.cpp:208-211
          if (tp != 3) { error = 13; return {}; }  // SEAL: not enough relinearization keys
          relinearize_inplace(c, temp.data(), relin_key);
          temp.resize(L);
          tp = 2;
        }
      }
      if (result.empty()) {
        result = temp;
        rp = tp;
      } else {
        for (size_t p = 0; p < rp; ++p)
          for (size_t j = 0; j < k; ++j) {
            const Modulus& m = c.mod(j);
            u64* x = result.data() + (p * k + j) * N;
            const u64* y = temp.data() + (p * k + j) * N;
            for (size_t n = 0; n < N; ++n) x[n] = addmod(x[n], y[n], m);
          }
      }
    }
    *polys = rp;
    return result;
  }
};

// PIRDatabase::multiply in ciphertext-multiplication mode: out = one ciphertext of *polys polynomials
static inline int db_multiply_ct(const Context& c, const RnsTool& R, const u64* db, size_t num_pt, const uint32_t* dims,
                                 size_t nd, const u64* sv, size_t n_sv, const u64* relin_key, std::vector<u64>& out,
                                 size_t* polys) {
  size_t dim_sum = 0;
  for (size_t i = 0; i < nd; ++i) dim_sum += dims[i];
  if (n_sv != dim_sum) return 3;
  CtDbMultiplier m(c, R, db, num_pt, sv, relin_key);
  out = m.multiply(dims, nd, 0, polys);
  return m.error;
}

// server.cpp:173-195 processQuery in ciphertext-multiplication mode
static inline int process_query_ct(const Context& c, const RnsTool& R, const u64* db, size_t num_pt, const uint32_t* dims,
                                   size_t nd, const GaloisKeys& gk, const u64* relin_key, const u64* query, size_t n_ct,
                                   std::vector<u64>& out, size_t* polys) {
  size_t dim_sum = 0;
  for (size_t i = 0; i < nd; ++i) dim_sum += dims[i];
  std::vector<u64> sv;
  int rc = oblivious_expansion_multi(c, query, n_ct, dim_sum, gk, sv);
  if (rc) return rc;
  return db_multiply_ct(c, R, db, num_pt, dims, nd, sv.data(), dim_sum, relin_key, out, polys);
}

}  // namespace orc
