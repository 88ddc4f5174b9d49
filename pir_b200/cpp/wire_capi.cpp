// wire_capi.cpp — plain-C entry points over wire.hpp so that the wire codec can be exercised from the python tests
// (tests/test_wire_cpu.py compares it with the python protobuf runtime).  Host-only; builds to
// pir_b200/lib/libpirb_wire.so with g++.  Output buffers are malloc'ed and released with pirw_free.
#include <cstdlib>
#include <cstring>

#include "wire.hpp"

using namespace pir::wire;

namespace {
int give(const std::string& s, uint8_t** out, size_t* out_len) {
  *out = (uint8_t*)std::malloc(s.size() ? s.size() : 1);
  if (!*out) return 13;
  std::memcpy(*out, s.data(), s.size());
  *out_len = s.size();
  return 0;
}
thread_local std::string g_err;
SealParams make_params(uint32_t N, const uint64_t* moduli, uint32_t n_moduli, uint64_t t) {
  SealParams p;
  p.poly_modulus_degree = N;
  p.coeff_modulus.assign(moduli, moduli + n_moduli);
  p.plain_modulus = t;
  return p;
}
}  // namespace

extern "C" {

void pirw_free(void* p) { std::free(p); }
const char* pirw_last_error() { return g_err.c_str(); }

void pirw_blake2b(uint8_t* out, size_t outlen, const uint8_t* in, size_t inlen, const uint8_t* key, size_t keylen) {
  blake::blake2b(out, outlen, in, inlen, key, keylen);
}
void pirw_blake2xb(uint8_t* out, size_t outlen, const uint8_t* in, size_t inlen, const uint8_t* key, size_t keylen) {
  blake::blake2xb(out, outlen, in, inlen, key, keylen);
}
void pirw_sample_poly_uniform(const uint64_t* seed8, uint32_t N, const uint64_t* moduli, uint32_t n_moduli,
                              uint64_t* dst) {
  seed_type s;
  std::memcpy(s.data(), seed8, sizeof(s));
  BlakePRNG prng(s);
  sample_poly_uniform(prng, N, moduli, n_moduli, dst);
}
void pirw_parms_id(uint32_t N, const uint64_t* moduli, uint32_t n_moduli, uint64_t t, uint64_t* id4) {
  const parms_id_type id = compute_parms_id(N, moduli, n_moduli, t);
  std::memcpy(id4, id.data(), 32);
}

// kind: 0 Ciphertexts, 1 Request, 2 Response, 3 PIRParameters.  Parses `in` and serializes it again.
int pirw_proto_roundtrip(int kind, const uint8_t* in, size_t in_len, uint8_t** out, size_t* out_len) {
  const std::string_view sv((const char*)in, in_len);
  bool ok = false;
  std::string s;
  if (kind == 0) { CiphertextsMsg m; ok = Parse(sv, &m); if (ok) s = Serialize(m); }
  else if (kind == 1) { RequestMsg m; ok = Parse(sv, &m); if (ok) s = Serialize(m); }
  else if (kind == 2) { ResponseMsg m; ok = Parse(sv, &m); if (ok) s = Serialize(m); }
  else if (kind == 3) { PIRParametersMsg m; ok = Parse(sv, &m); if (ok) s = Serialize(m); }
  if (!ok) { g_err = "malformed protobuf message"; return 3; }
  return give(s, out, out_len);
}
// Builds a Request from parts: n_queries x cts_per_query blobs of equal length ct_len laid out back to back.
int pirw_request_build(const uint8_t* cts, uint32_t n_queries, uint32_t cts_per_query, size_t ct_len,
                       const uint8_t* gk, size_t gk_len, const uint8_t* rk, size_t rk_len, uint8_t** out,
                       size_t* out_len) {
  RequestMsg m;
  for (uint32_t q = 0; q < n_queries; ++q) {
    m.query.emplace_back();
    for (uint32_t c = 0; c < cts_per_query; ++c)
      m.query.back().ct.emplace_back((const char*)cts + ((size_t)q * cts_per_query + c) * ct_len, ct_len);
  }
  m.galois_keys.assign((const char*)gk, gk_len);
  m.relin_keys.assign((const char*)rk, rk_len);
  return give(Serialize(m), out, out_len);
}
int pirw_params_build(uint64_t num_items, uint64_t num_pt, const uint32_t* dims, uint32_t n_dims, const uint8_t* ep,
                      size_t ep_len, uint32_t bytes_per_item, uint32_t items_per_pt, uint32_t bits_per_coeff,
                      int use_ct_mult, uint8_t** out, size_t* out_len) {
  PIRParametersMsg m;
  m.num_items = num_items;
  m.num_pt = num_pt;
  m.dimensions.assign(dims, dims + n_dims);
  m.encryption_parameters.assign((const char*)ep, ep_len);
  m.bytes_per_item = bytes_per_item;
  m.items_per_plaintext = items_per_pt;
  m.bits_per_coeff = bits_per_coeff;
  m.use_ciphertext_multiplication = use_ct_mult != 0;
  return give(Serialize(m), out, out_len);
}

// ---- SEAL objects ----
int pirw_ct_save(const uint64_t* limbs, uint32_t size, uint32_t N, uint32_t n_moduli, const uint64_t* parms_id4,
                 int is_ntt, const uint64_t* seed8_or_null, uint8_t** out, size_t* out_len) {
  CiphertextData ct;
  std::memcpy(ct.parms_id.data(), parms_id4, 32);
  ct.is_ntt_form = is_ntt != 0;
  ct.size = size;
  ct.poly_modulus_degree = N;
  ct.coeff_modulus_size = n_moduli;
  ct.limbs.assign(limbs, limbs + (size_t)size * n_moduli * N);
  seed_type s;
  if (seed8_or_null) std::memcpy(s.data(), seed8_or_null, sizeof(s));
  return give(SaveCiphertext(ct, seed8_or_null ? &s : nullptr), out, out_len);
}
// limbs_out must hold 2 * n_moduli * N words (size-2 ciphertexts only)
int pirw_ct_load(const uint8_t* in, size_t in_len, uint32_t N, const uint64_t* moduli, uint32_t n_moduli,
                 uint64_t* limbs_out, uint64_t* parms_id4, int* is_ntt, int* was_seeded) {
  CiphertextData ct;
  if (!LoadCiphertext(std::string_view((const char*)in, in_len), N, moduli, n_moduli, &ct, &g_err)) return 3;
  if (ct.size != 2) { g_err = "only size-2 ciphertexts are supported"; return 3; }
  std::memcpy(limbs_out, ct.limbs.data(), ct.limbs.size() * 8);
  std::memcpy(parms_id4, ct.parms_id.data(), 32);
  *is_ntt = ct.is_ntt_form;
  *was_seeded = ct.was_seeded;
  return 0;
}
// Same for a ciphertext of up to cap_polys polynomials (ciphertext-multiplication replies without relinearization
// keep a third polynomial, server.cpp:185-190); *size_out receives the count.
int pirw_ct_load_any(const uint8_t* in, size_t in_len, uint32_t N, const uint64_t* moduli, uint32_t n_moduli,
                     uint64_t* limbs_out, uint32_t cap_polys, uint32_t* size_out, uint64_t* parms_id4, int* is_ntt,
                     int* was_seeded) {
  CiphertextData ct;
  if (!LoadCiphertext(std::string_view((const char*)in, in_len), N, moduli, n_moduli, &ct, &g_err)) return 3;
  if (ct.size > cap_polys) { g_err = "ciphertext larger than expected"; return 3; }
  std::memcpy(limbs_out, ct.limbs.data(), ct.limbs.size() * 8);
  std::memcpy(parms_id4, ct.parms_id.data(), 32);
  *size_out = (uint32_t)ct.size;
  *is_ntt = ct.is_ntt_form;
  *was_seeded = ct.was_seeded;
  return 0;
}
// Galois keys in the C-ABI layout: elts[n], limbs [n][k][2][k+1][N].  seeds (optional) [n][k][8]: write every key
// seed-compressed, its second polynomial REPLACED by the expansion of its seed (as SEAL's keygen produces them).
int pirw_galois_keys_save(uint32_t N, const uint64_t* moduli, uint32_t n_moduli, uint64_t t, const uint32_t* elts,
                          uint32_t n, const uint64_t* limbs, const uint64_t* seeds_or_null, uint8_t** out,
                          size_t* out_len) {
  const SealParams P = make_params(N, moduli, n_moduli, t);
  const uint32_t k = n_moduli - 1;
  KSwitchKeysData K;
  K.parms_id = key_parms_id(P);
  K.keys.assign(N, {});  // KeyGenerator::galois_keys sizes the table to poly_modulus_degree slots
  std::vector<std::vector<seed_type>> seeds(N);
  const size_t ctw = (size_t)2 * n_moduli * N;
  for (uint32_t e = 0; e < n; ++e) {
    const uint32_t idx = galois_index(elts[e]);
    if (idx >= N) { g_err = "Galois element out of range"; return 3; }
    K.keys[idx].resize(k);
    seeds[idx].resize(k);
    for (uint32_t j = 0; j < k; ++j) {
      CiphertextData& ct = K.keys[idx][j];
      ct.parms_id = K.parms_id;
      ct.is_ntt_form = true;
      ct.size = 2;
      ct.poly_modulus_degree = N;
      ct.coeff_modulus_size = n_moduli;
      const uint64_t* src = limbs + ((size_t)e * k + j) * ctw;
      ct.limbs.assign(src, src + ctw);
      if (seeds_or_null) std::memcpy(seeds[idx][j].data(), seeds_or_null + ((size_t)e * k + j) * 8, sizeof(seed_type));
    }
  }
  return give(SaveKSwitchKeys(K, seeds_or_null ? &seeds : nullptr), out, out_len);
}
// Returns the number of keys through *n_out; elts_out[max_n], limbs_out [max_n][k][2][k+1][N]
int pirw_galois_keys_load(const uint8_t* in, size_t in_len, uint32_t N, const uint64_t* moduli, uint32_t n_moduli,
                          uint64_t t, uint32_t max_n, uint32_t* elts_out, uint64_t* limbs_out, uint32_t* n_out) {
  const SealParams P = make_params(N, moduli, n_moduli, t);
  KSwitchKeysData K;
  if (!LoadKSwitchKeys(std::string_view((const char*)in, in_len), P, &K, &g_err)) return 3;
  const uint32_t k = n_moduli - 1;
  const size_t ctw = (size_t)2 * n_moduli * N;
  uint32_t n = 0;
  for (size_t s = 0; s < K.keys.size(); ++s) {
    if (K.keys[s].empty()) continue;
    if (n >= max_n) { g_err = "too many keys"; return 3; }
    elts_out[n] = galois_elt_of_index((uint32_t)s);
    for (uint32_t j = 0; j < k; ++j)
      std::memcpy(limbs_out + ((size_t)n * k + j) * ctw, K.keys[s][j].limbs.data(), ctw * 8);
    ++n;
  }
  *n_out = n;
  return 0;
}
int pirw_encryption_parameters_save(uint32_t N, const uint64_t* moduli, uint32_t n_moduli, uint64_t t, uint8_t** out,
                                    size_t* out_len) {
  return give(SaveEncryptionParameters(make_params(N, moduli, n_moduli, t)), out, out_len);
}
int pirw_encryption_parameters_load(const uint8_t* in, size_t in_len, uint32_t* N, uint64_t* moduli, uint32_t max_moduli,
                                    uint32_t* n_moduli, uint64_t* t) {
  SealParams p;
  if (!LoadEncryptionParameters(std::string_view((const char*)in, in_len), &p, &g_err)) return 3;
  if (p.coeff_modulus.size() > max_moduli) { g_err = "too many moduli"; return 3; }
  *N = p.poly_modulus_degree;
  *n_moduli = (uint32_t)p.coeff_modulus.size();
  std::memcpy(moduli, p.coeff_modulus.data(), p.coeff_modulus.size() * 8);
  *t = p.plain_modulus;
  return 0;
}

}  // extern "C"

// ---- message walking for host languages without a protobuf runtime (pir_b200/wire.py) ----
namespace {
struct ParsedMsg {
  std::vector<CiphertextsMsg> groups;  // Request.query or Response.reply
  std::string galois_keys, relin_keys;
};
}  // namespace

extern "C" {

// kind: 1 Request, 2 Response.  Returns an opaque handle through *out (release with pirw_msg_free).
int pirw_msg_parse(int kind, const uint8_t* in, size_t in_len, void** out) {
  auto* m = new ParsedMsg();
  bool ok = false;
  const std::string_view sv((const char*)in, in_len);
  if (kind == 1) {
    RequestMsg r;
    ok = Parse(sv, &r);
    if (ok) { m->groups = std::move(r.query); m->galois_keys = std::move(r.galois_keys); m->relin_keys = std::move(r.relin_keys); }
  } else if (kind == 2) {
    ResponseMsg r;
    ok = Parse(sv, &r);
    if (ok) m->groups = std::move(r.reply);
  }
  if (!ok) { delete m; g_err = "malformed protobuf message"; return 3; }
  *out = m;
  return 0;
}
void pirw_msg_free(void* h) { delete (ParsedMsg*)h; }
uint32_t pirw_msg_groups(const void* h) { return (uint32_t)((const ParsedMsg*)h)->groups.size(); }
uint32_t pirw_msg_group_size(const void* h, uint32_t g) {
  const auto* m = (const ParsedMsg*)h;
  return g < m->groups.size() ? (uint32_t)m->groups[g].ct.size() : 0;
}
// Borrowed pointers, valid until pirw_msg_free.
int pirw_msg_ct(const void* h, uint32_t g, uint32_t i, const uint8_t** p, size_t* len) {
  const auto* m = (const ParsedMsg*)h;
  if (g >= m->groups.size() || i >= m->groups[g].ct.size()) { g_err = "index out of range"; return 3; }
  *p = (const uint8_t*)m->groups[g].ct[i].data();
  *len = m->groups[g].ct[i].size();
  return 0;
}
int pirw_msg_keys(const void* h, int which, const uint8_t** p, size_t* len) {  // which: 2 galois_keys, 3 relin_keys
  const auto* m = (const ParsedMsg*)h;
  const std::string& s = which == 2 ? m->galois_keys : m->relin_keys;
  *p = (const uint8_t*)s.data();
  *len = s.size();
  return 0;
}
// Response from n_groups x per_group serialized ciphertexts of equal length laid out back to back.
int pirw_response_build(const uint8_t* cts, uint32_t n_groups, uint32_t per_group, size_t ct_len, uint8_t** out,
                        size_t* out_len) {
  ResponseMsg m;
  for (uint32_t q = 0; q < n_groups; ++q) {
    m.reply.emplace_back();
    for (uint32_t c = 0; c < per_group; ++c)
      m.reply.back().ct.emplace_back((const char*)cts + ((size_t)q * per_group + c) * ct_len, ct_len);
  }
  return give(Serialize(m), out, out_len);
}
// Structural check of a KSwitchKeys blob without keeping its data (relin_keys are parsed and otherwise unused).
int pirw_kswitch_keys_check(const uint8_t* in, size_t in_len, uint32_t N, const uint64_t* moduli, uint32_t n_moduli,
                            uint64_t t) {
  KSwitchKeysData K;
  return LoadKSwitchKeys(std::string_view((const char*)in, in_len), make_params(N, moduli, n_moduli, t), &K, &g_err,
                         false) ? 0 : 3;
}

}  // extern "C"
