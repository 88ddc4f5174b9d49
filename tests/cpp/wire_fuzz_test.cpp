// CPU-only robustness test of the wire layer (pir_b200/cpp/wire.hpp), built with AddressSanitizer + UBSan: the server
// parses bytes from the network (pir::PIRServer::ProcessRequest(const std::string&)), so truncated, bit-flipped and
// length-inflated requests must be rejected — or parsed — without reading or writing out of bounds, and the copy-free
// parsers the device server uses (ParseView, LoadCiphertextTo) must agree with the object-building ones on every input.
#include <cstdio>
#include <cstring>
#include <random>

#include "../../pir_b200/cpp/wire.hpp"

#define CHECK(cond, msg)                                                \
  do {                                                                  \
    if (!(cond)) {                                                      \
      std::fprintf(stderr, "FAIL %s:%d %s\n", __FILE__, __LINE__, msg); \
      return 1;                                                         \
    }                                                                   \
  } while (0)

namespace w = pir::wire;

int main(int argc, char** argv) {
  const int iterations = argc > 1 ? std::atoi(argv[1]) : 1500;
  const uint32_t N = 4096;
  w::SealParams sp;
  sp.poly_modulus_degree = N;
  sp.coeff_modulus = {0xffffee001ULL, 0xffffc4001ULL, 0x1ffffe0001ULL};
  sp.plain_modulus = 0xFC001;
  const size_t k = 2;
  std::mt19937_64 rng(77);
  auto ct_of = [&](size_t polys, size_t n_mod, bool ntt, const w::parms_id_type& pid) {
    w::CiphertextData c;
    c.parms_id = pid;
    c.is_ntt_form = ntt;
    c.size = polys;
    c.poly_modulus_degree = N;
    c.coeff_modulus_size = n_mod;
    c.limbs.resize(polys * n_mod * N);
    for (size_t p = 0; p < polys; ++p)
      for (size_t j = 0; j < n_mod; ++j)
        for (size_t n = 0; n < N; ++n) c.limbs[(p * n_mod + j) * N + n] = rng() % sp.coeff_modulus[j];
    return c;
  };
  // a well-formed request: 2 queries x 1 ciphertext (one of them seed-compressed), 2 Galois keys, 1 relinearization key
  w::RequestMsg req;
  w::seed_type seed{};
  for (auto& x : seed) x = rng();
  for (int q = 0; q < 2; ++q) {
    req.query.emplace_back();
    req.query.back().ct.push_back(w::SaveCiphertext(ct_of(2, k, false, w::data_parms_id(sp)), q ? &seed : nullptr));
  }
  auto keys_of = [&](std::vector<uint32_t> slots, size_t n_slots) {
    w::KSwitchKeysData K;
    K.parms_id = w::key_parms_id(sp);
    K.keys.assign(n_slots, {});
    for (uint32_t s : slots)
      for (size_t j = 0; j < k; ++j) K.keys[s].push_back(ct_of(2, k + 1, true, K.parms_id));
    return w::SaveKSwitchKeys(K);
  };
  req.galois_keys = keys_of({w::galois_index(3), w::galois_index(N + 1)}, N);
  req.relin_keys = keys_of({0}, 1);
  const std::string good = w::Serialize(req);
  std::printf("request: %zu bytes\n", good.size());

  auto exercise = [&](const std::string& bytes, bool must_parse) -> int {
    w::RequestMsg m;
    w::RequestView v;
    const bool ok1 = w::Parse(bytes, &m), ok2 = w::ParseView(bytes, &v);
    CHECK(ok1 == ok2, "Parse and ParseView disagree on validity");
    if (must_parse) CHECK(ok1, "the well-formed request must parse");
    if (!ok1) return 0;
    CHECK(m.query.size() == v.query.size() && std::string_view(m.galois_keys) == v.galois_keys &&
              std::string_view(m.relin_keys) == v.relin_keys, "Parse and ParseView disagree on content");
    std::string err;
    for (size_t q = 0; q < m.query.size(); ++q) {
      CHECK(m.query[q].ct.size() == v.query[q].size(), "query shape");
      for (size_t c = 0; c < v.query[q].size(); ++c) {
        CHECK(std::string_view(m.query[q].ct[c]) == v.query[q][c], "ciphertext bytes");
        w::CiphertextData obj, meta;
        std::vector<uint64_t> direct(2 * k * N, ~0ull);
        const bool a = w::LoadCiphertext(v.query[q][c], N, sp.coeff_modulus.data(), k, &obj, &err);
        const bool b = w::LoadCiphertextTo(v.query[q][c], N, sp.coeff_modulus.data(), k, direct.data(), &meta, &err);
        CHECK(!b || a, "LoadCiphertextTo accepted what LoadCiphertext refuses");
        if (a && obj.size == 2) CHECK(b && direct == obj.limbs, "in-place load differs from the object load");
        if (must_parse) CHECK(a && b, "well-formed ciphertext refused");
      }
    }
    w::KSwitchKeysData G, R;
    const bool g = w::LoadKSwitchKeys(v.galois_keys, sp, &G, &err);
    const bool r = w::LoadKSwitchKeys(v.relin_keys, sp, &R, &err, /*keep_data=*/false);
    if (must_parse) CHECK(g && r && G.keys.size() == N && R.keys.size() == 1, "well-formed keys refused");
    return 0;
  };
  if (exercise(good, true)) return 1;

  int parsed = 0;
  for (int it = 0; it < iterations; ++it) {
    std::string bad = good;
    switch (it % 5) {
      case 0:  // truncation, mostly inside the small structures at the front and at field boundaries
        bad.resize(it % 10 == 0 ? rng() % good.size() : rng() % 400000);
        break;
      case 1:  // single bit flip in the first 300 KB (headers, lengths, the query ciphertexts)
        bad[rng() % 300000] ^= (char)(1u << (rng() % 8));
        break;
      case 2: {  // a window of 0xFF: inflated varints / sizes
        const size_t at = rng() % (good.size() - 16);
        std::memset(&bad[at], 0xFF, 1 + rng() % 10);
        break;
      }
      case 3: {  // random garbage over a header-sized window anywhere
        const size_t at = rng() % (good.size() - 32);
        for (size_t i = 0; i < 24; ++i) bad[at + i] = (char)rng();
        break;
      }
      default: {  // splice: a prefix of the request followed by a copy of its own beginning
        const size_t cut = rng() % 300000;
        bad = good.substr(0, cut) + good.substr(0, rng() % 5000);
        break;
      }
    }
    w::RequestMsg probe;
    parsed += w::Parse(bad, &probe) ? 1 : 0;
    if (exercise(bad, false)) return 1;
  }
  // the parameters message and bare objects on garbage
  for (int it = 0; it < 2000; ++it) {
    std::string junk(rng() % 200, 0);
    for (auto& ch : junk) ch = (char)rng();
    w::PIRParametersMsg pm;
    w::ResponseMsg rm;
    w::SealParams p2;
    w::CiphertextData c;
    w::KSwitchKeysData K;
    std::string err;
    (void)w::Parse(junk, &pm);
    (void)w::Parse(junk, &rm);
    (void)w::LoadEncryptionParameters(junk, &p2, &err);
    (void)w::LoadCiphertext(junk, N, sp.coeff_modulus.data(), k, &c, &err);
    (void)w::LoadKSwitchKeys(junk, sp, &K, &err);
  }
  std::printf("WIRE_FUZZ_TEST_OK: %d mutated requests (%d still framed as protobuf), no sanitizer report\n", iterations, parsed);
  return 0;
}
