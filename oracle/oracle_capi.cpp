// oracle/oracle_capi.cpp — C ABI over the CPU oracle (TEST INFRASTRUCTURE ONLY; see pir_oracle.hpp).
// Loaded with ctypes by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
#include <memory>

#include "pir_oracle.hpp"

using namespace orc;

struct OrcHandle {
  Context ctx;
  Crypto crypto;
  std::unique_ptr<RnsTool> rns_;  // built on first use (ciphertext-multiplication mode only)
  OrcHandle(size_t N, const std::vector<u64>& mods, u64 t) : ctx(N, mods, t), crypto(ctx) {}
  const RnsTool& rns() {
    if (!rns_) rns_.reset(new RnsTool(ctx));
    return *rns_;
  }
};

static GaloisKeys make_keys(const OrcHandle* h, const uint32_t* elts, size_t n_elts, const u64* limbs) {
  GaloisKeys gk;
  gk.elts.assign(elts, elts + n_elts);
  size_t per = gk.key_limbs(h->ctx);
  gk.data.assign(limbs, limbs + per * n_elts);
  return gk;
}

extern "C" {

// ---- context -----------------------------------------------------------------
void* orc_create(uint64_t N, uint32_t n_moduli, const uint64_t* moduli, uint64_t t) {
  try {
    return new OrcHandle(N, std::vector<u64>(moduli, moduli + n_moduli), t);
  } catch (...) {
    return nullptr;
  }
}
void orc_destroy(void* h) { delete (OrcHandle*)h; }
uint32_t orc_expansion_ratio(void* h) { return ((OrcHandle*)h)->ctx.expansion_ratio(); }
uint64_t orc_psi(void* h, uint32_t j) { return ((OrcHandle*)h)->ctx.tb[j].psi; }

// ---- parameter math ------------------------------------------------------------
uint64_t orc_plain_modulus_batching(uint64_t N, int bits) {
  try { return plain_modulus_batching(N, bits); } catch (...) { return 0; }
}
int orc_bfv_default(uint64_t N, uint64_t* out, uint32_t cap) {
  try {
    auto v = bfv_default_coeff_modulus(N);
    if (v.size() > cap) return -1;
    for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
    return (int)v.size();
  } catch (...) { return -1; }
}
int orc_is_prime(uint64_t v) { return is_prime_u64(v) ? 1 : 0; }
void orc_calculate_dimensions(uint32_t db_size, uint32_t nd, uint32_t* out) {
  auto v = calculate_dimensions(db_size, nd);
  for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
}
uint64_t orc_next_power_two(uint64_t v) { return next_power_two<size_t>(v); }
uint32_t orc_ceil_log2(uint32_t v) { return ceil_log2_u32(v); }
uint32_t orc_log2(uint32_t v) { return log2_u32(v); }

// ---- ring primitives -----------------------------------------------------------
void orc_ntt_forward(void* h, uint32_t j, uint64_t* poly) { ((OrcHandle*)h)->ctx.tb[j].forward(poly); }
void orc_ntt_inverse(void* h, uint32_t j, uint64_t* poly) { ((OrcHandle*)h)->ctx.tb[j].inverse(poly); }
uint64_t orc_mulmod(void* h, uint32_t j, uint64_t a, uint64_t b) { return mulmod(a, b, ((OrcHandle*)h)->ctx.mod(j)); }
void orc_plain_to_ntt(void* h, const uint64_t* coeffs, uint64_t n_coeff, uint64_t* out) {
  plain_to_ntt(((OrcHandle*)h)->ctx, coeffs, n_coeff, out);
}
void orc_ct_to_ntt(void* h, uint64_t* ct) { ct_to_ntt(((OrcHandle*)h)->ctx, ct); }
void orc_ct_from_ntt(void* h, uint64_t* ct) { ct_from_ntt(((OrcHandle*)h)->ctx, ct); }

// ---- server path ---------------------------------------------------------------
int orc_substitute(void* hh, uint64_t* ct, uint32_t power, const uint32_t* elts, uint32_t n_elts, const uint64_t* keys) {
  OrcHandle* h = (OrcHandle*)hh;
  GaloisKeys gk = make_keys(h, elts, n_elts, keys);
  return apply_galois_inplace(h->ctx, ct, power, gk) ? 0 : 13;
}
void orc_mul_inv_pow_x(void* hh, const uint64_t* in, uint32_t kpow, uint64_t* out) {
  multiply_inverse_power_of_x(((OrcHandle*)hh)->ctx, in, kpow, out);
}
// out must hold total_items cts
int orc_expand(void* hh, const uint64_t* cts, uint64_t n_ct, uint64_t total_items, const uint32_t* elts, uint32_t n_elts,
               const uint64_t* keys, uint64_t* out, int single) {
  OrcHandle* h = (OrcHandle*)hh;
  GaloisKeys gk = make_keys(h, elts, n_elts, keys);
  std::vector<u64> v;
  int rc = single ? oblivious_expansion(h->ctx, cts, total_items, gk, v)
                  : oblivious_expansion_multi(h->ctx, cts, n_ct, total_items, gk, v);
  if (rc) return rc;
  std::memcpy(out, v.data(), v.size() * sizeof(u64));
  return 0;
}
void orc_reencode(void* hh, const uint64_t* ct, uint64_t* pts) {
  std::vector<u64> v;
  reencode(((OrcHandle*)hh)->ctx, ct, v);
  std::memcpy(pts, v.data(), v.size() * sizeof(u64));
}
void orc_reencode_decode(void* hh, const uint64_t* pts, uint64_t* ct) {
  OrcHandle* h = (OrcHandle*)hh;
  std::memset(ct, 0, h->ctx.ct_limbs() * sizeof(u64));
  reencode_decode(h->ctx, pts, ct);
}
// sv is mutated (NTT form on exit for the entries that were used). out_count receives #cts.
int orc_db_multiply(void* hh, const uint64_t* db, uint64_t num_pt, const uint32_t* dims, uint32_t nd, uint64_t* sv,
                    uint64_t n_sv, uint64_t* out, uint64_t out_cap_cts, uint64_t* out_count) {
  OrcHandle* h = (OrcHandle*)hh;
  std::vector<u64> v;
  int rc = db_multiply(h->ctx, db, num_pt, dims, nd, sv, n_sv, v);
  if (rc) return rc;
  size_t n = v.size() / h->ctx.ct_limbs();
  if (n > out_cap_cts) return 13;
  std::memcpy(out, v.data(), v.size() * sizeof(u64));
  *out_count = n;
  return 0;
}
int orc_process_query(void* hh, const uint64_t* db, uint64_t num_pt, const uint32_t* dims, uint32_t nd,
                      const uint32_t* elts, uint32_t n_elts, const uint64_t* keys, const uint64_t* query, uint64_t n_ct,
                      uint64_t* out, uint64_t out_cap_cts, uint64_t* out_count) {
  OrcHandle* h = (OrcHandle*)hh;
  GaloisKeys gk = make_keys(h, elts, n_elts, keys);
  std::vector<u64> v;
  int rc = process_query(h->ctx, db, num_pt, dims, nd, gk, query, n_ct, v);
  if (rc) return rc;
  size_t n = v.size() / h->ctx.ct_limbs();
  if (n > out_cap_cts) return 13;
  std::memcpy(out, v.data(), v.size() * sizeof(u64));
  *out_count = n;
  return 0;
}
// scan only (last-dimension inner product of one row against `count` NTT-form plaintexts), for CPU GB/s timing.
// sv_ntt: [count][2][k][N] NTT form; out: one ct in NTT form.
void orc_scan_row(void* hh, const uint64_t* db, uint64_t count, const uint64_t* sv_ntt, uint64_t* out) {
  OrcHandle* h = (OrcHandle*)hh;
  const Context& c = h->ctx;
  std::vector<u64> tmp(c.ct_limbs());
  for (u64 i = 0; i < count; ++i) {
    if (i == 0) multiply_plain_ntt(c, sv_ntt, db, out);
    else {
      multiply_plain_ntt(c, sv_ntt + i * c.ct_limbs(), db + i * c.pt_limbs(), tmp.data());
      add_ct_inplace(c, out, tmp.data());
    }
  }
}

// ---- string encoder ------------------------------------------------------------
int64_t orc_string_encode(const uint8_t* bytes, uint64_t n, uint64_t bits, uint64_t* coeffs, uint64_t max_coeff) {
  size_t r = string_encode(bytes, n, bits, coeffs, max_coeff);
  return r == (size_t)-1 ? -1 : (int64_t)r;
}
int orc_string_decode(const uint64_t* coeffs, uint64_t n_coeff, uint64_t bits, uint64_t length, uint64_t offset, uint8_t* out) {
  return string_decode(coeffs, n_coeff, bits, length, offset, out);
}

// ---- harness crypto ------------------------------------------------------------
// sk_ntt: [k+1][N]; sk_coeff: int8 [N]
void orc_keygen(void* hh, uint64_t seed, uint64_t* sk_ntt, int8_t* sk_coeff, uint64_t* pk_ntt) {
  OrcHandle* h = (OrcHandle*)hh;
  Rng rng(seed);
  SecretKey sk = gen_secret_key(h->ctx, rng);
  PublicKey pk = gen_public_key(h->ctx, sk, rng);
  std::memcpy(sk_ntt, sk.ntt.data(), sk.ntt.size() * sizeof(u64));
  std::memcpy(sk_coeff, sk.coeff.data(), sk.coeff.size());
  std::memcpy(pk_ntt, pk.ntt.data(), pk.ntt.size() * sizeof(u64));
}
void orc_gen_galois_keys(void* hh, uint64_t seed, const uint64_t* sk_ntt, const int8_t* sk_coeff, const uint32_t* elts,
                         uint32_t n_elts, uint64_t* out) {
  OrcHandle* h = (OrcHandle*)hh;
  const Context& c = h->ctx;
  SecretKey sk;
  sk.ntt.assign(sk_ntt, sk_ntt + (c.k + 1) * c.N);
  sk.coeff.assign(sk_coeff, sk_coeff + c.N);
  Rng rng(seed);
  size_t per = c.k * 2 * (c.k + 1) * c.N;
  for (uint32_t i = 0; i < n_elts; ++i) gen_galois_key(c, sk, elts[i], rng, out + i * per);
}
void orc_encrypt(void* hh, uint64_t seed, const uint64_t* pk_ntt, const uint64_t* pt, uint64_t n_coeff, uint64_t* ct) {
  OrcHandle* h = (OrcHandle*)hh;
  PublicKey pk;
  pk.ntt.assign(pk_ntt, pk_ntt + 2 * (h->ctx.k + 1) * h->ctx.N);
  Rng rng(seed);
  h->crypto.encrypt(pk, pt, n_coeff, rng, ct);
}
// ---- ciphertext-multiplication mode (bfv_mul_oracle.hpp) -----------------------------
// primes: m_sk, then the nB primes of B.  Returns nB.
uint32_t orc_rns_bases(void* hh, uint64_t* primes, uint32_t cap) {
  OrcHandle* h = (OrcHandle*)hh;
  const RnsTool& R = h->rns();
  if (cap >= R.nB + 1) {
    primes[0] = R.m_sk;
    for (size_t i = 0; i < R.nB; ++i) primes[1 + i] = R.bsk[i].mod.q;
  }
  return (uint32_t)R.nB;
}
// out: s1 + s2 - 1 polynomials [.][k][N]
void orc_bfv_multiply(void* hh, const uint64_t* a, uint32_t s1, const uint64_t* b, uint32_t s2, uint64_t* out) {
  OrcHandle* h = (OrcHandle*)hh;
  std::vector<u64> v = bfv_multiply(h->ctx, h->rns(), a, s1, b, s2);
  std::memcpy(out, v.data(), v.size() * sizeof(u64));
}
// ct: [3][k][N]; the first two polynomials are updated
void orc_relinearize(void* hh, uint64_t* ct, const uint64_t* relin_key) { relinearize_inplace(((OrcHandle*)hh)->ctx, ct, relin_key); }
void orc_gen_relin_key(void* hh, uint64_t seed, const uint64_t* sk_ntt, const int8_t* sk_coeff, uint64_t* out) {
  OrcHandle* h = (OrcHandle*)hh;
  const Context& c = h->ctx;
  SecretKey sk;
  sk.ntt.assign(sk_ntt, sk_ntt + (c.k + 1) * c.N);
  sk.coeff.assign(sk_coeff, sk_coeff + c.N);
  Rng rng(seed);
  gen_relin_key(c, sk, rng, out);
}
int orc_decrypt_polys(void* hh, const uint64_t* sk_ntt, const uint64_t* ct, uint32_t polys, uint64_t* pt) {
  OrcHandle* h = (OrcHandle*)hh;
  SecretKey sk;
  sk.ntt.assign(sk_ntt, sk_ntt + (h->ctx.k + 1) * h->ctx.N);
  return decrypt_any(h->crypto, sk, ct, polys, pt);
}
// relin_key may be null.  out: one ciphertext of *out_polys polynomials (at most out_cap_polys).
int orc_db_multiply_ct(void* hh, const uint64_t* db, uint64_t num_pt, const uint32_t* dims, uint32_t nd, const uint64_t* sv,
                       uint64_t n_sv, const uint64_t* relin_key, uint64_t* out, uint32_t out_cap_polys, uint32_t* out_polys) {
  OrcHandle* h = (OrcHandle*)hh;
  std::vector<u64> v;
  size_t polys = 0;
  int rc = db_multiply_ct(h->ctx, h->rns(), db, num_pt, dims, nd, sv, n_sv, relin_key, v, &polys);
  if (rc) return rc;
  if (polys > out_cap_polys) return 13;
  std::memcpy(out, v.data(), v.size() * sizeof(u64));
  *out_polys = (uint32_t)polys;
  return 0;
}
int orc_process_query_ct(void* hh, const uint64_t* db, uint64_t num_pt, const uint32_t* dims, uint32_t nd,
                         const uint32_t* elts, uint32_t n_elts, const uint64_t* keys, const uint64_t* relin_key,
                         const uint64_t* query, uint64_t n_ct, uint64_t* out, uint32_t out_cap_polys, uint32_t* out_polys) {
  OrcHandle* h = (OrcHandle*)hh;
  GaloisKeys gk = make_keys(h, elts, n_elts, keys);
  std::vector<u64> v;
  size_t polys = 0;
  int rc = process_query_ct(h->ctx, h->rns(), db, num_pt, dims, nd, gk, relin_key, query, n_ct, v, &polys);
  if (rc) return rc;
  if (polys > out_cap_polys) return 13;
  std::memcpy(out, v.data(), v.size() * sizeof(u64));
  *out_polys = (uint32_t)polys;
  return 0;
}

int orc_decrypt(void* hh, const uint64_t* sk_ntt, const uint64_t* ct, uint64_t* pt) {
  OrcHandle* h = (OrcHandle*)hh;
  SecretKey sk;
  sk.ntt.assign(sk_ntt, sk_ntt + (h->ctx.k + 1) * h->ctx.N);
  return h->crypto.decrypt(sk, ct, pt);
}

}  // extern "C"
