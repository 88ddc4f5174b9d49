// Host-only tests of the C++ shim (pir_b200/cpp/pir_b200.hpp), written like the reference's own tests of the same
// functions: parameters_test.cpp:47-110, string_encoder_test.cpp:64-126, database_test.cpp:390-464.  No device needed:
// parameter math, string packing and index arithmetic are host code (the shape helpers come from libpirb200.so).
#include <cmath>
#include <cstdio>
#include <random>

#include "../../pir_b200/cpp/pir_b200.hpp"

#define CHECK(cond, msg)                                                \
  do {                                                                  \
    if (!(cond)) {                                                      \
      std::fprintf(stderr, "FAIL %s:%d %s\n", __FILE__, __LINE__, msg); \
      return 1;                                                         \
    }                                                                   \
  } while (0)

using U32 = std::vector<uint32_t>;

static int parameters_tests() {
  {  // SanityCheck (parameters_test.cpp:47-62)
    auto p = pir::CreatePIRParameters(1026, 256);
    CHECK(p.ok(), "CreatePIRParameters(1026, 256)");
    CHECK((*p)->num_items == 1026 && (*p)->num_pt == 27 && (*p)->bytes_per_item == 256 &&
              (*p)->items_per_plaintext == 38 && (*p)->dimensions == U32{27}, "SanityCheck values");
    CHECK((*p)->encryption_parameters.poly_modulus_degree == 4096 && (*p)->encryption_parameters.plain_modulus == 0xFC001 &&
              (*p)->encryption_parameters.coeff_modulus.size() == 3, "default encryption parameters");
  }
  {  // CreateMultiDim (:64-79)
    auto p = pir::CreatePIRParameters(19011, 500, 3);
    CHECK(p.ok() && (*p)->num_pt == 1001 && (*p)->items_per_plaintext == 19 && (*p)->dimensions == (U32{11, 10, 10}),
          "CreateMultiDim");
  }
  {  // CreateAllParams (:81-99)
    auto p = pir::CreatePIRParameters(77412, 777, 2, pir::GenerateEncryptionParams(8192), true, 12);
    CHECK(p.ok() && (*p)->num_pt == 5161 && (*p)->bytes_per_item == 777 && (*p)->items_per_plaintext == 15 &&
              (*p)->dimensions == (U32{72, 72}) && (*p)->bits_per_coeff == 12 && (*p)->use_ciphertext_multiplication,
          "CreateAllParams");
    auto plain = pir::CreatePIRParameters(77412, 777, 2, pir::GenerateEncryptionParams(8192), false, 12);
    CHECK(plain.ok() && !(*plain)->use_ciphertext_multiplication, "flag off by default");
    auto too_many_bits = pir::CreatePIRParameters(100, 0, 1, pir::GenerateEncryptionParams(4096), false, 25);
    CHECK(!too_many_bits.ok() && too_many_bits.status().code() == PIRB_INVALID_ARGUMENT, "bits_per_coeff above max");
    auto too_big = pir::CreatePIRParameters(100, 99999, 1);
    CHECK(!too_big.ok() && too_big.status().code() == PIRB_INVALID_ARGUMENT, "item larger than a plaintext");
  }
  {  // EncryptionParamsSerialization (:101-110) and the PIRParameters message as a whole
    auto ep = pir::GenerateEncryptionParams(8192);
    pir::wire::SealParams sp;
    std::string err;
    CHECK(pir::wire::LoadEncryptionParameters(pir::wire::SaveEncryptionParameters(pir::ToSealParams(ep)), &sp, &err),
          err.c_str());
    CHECK(sp.poly_modulus_degree == 8192 && sp.coeff_modulus == ep.coeff_modulus && sp.plain_modulus == ep.plain_modulus,
          "encryption parameters round trip");
    auto p = *pir::CreatePIRParameters(19011, 500, 3);
    auto q = pir::ParsePIRParameters(pir::SerializePIRParameters(*p));
    CHECK(q.ok() && (*q)->num_items == p->num_items && (*q)->num_pt == p->num_pt && (*q)->dimensions == p->dimensions &&
              (*q)->items_per_plaintext == p->items_per_plaintext &&
              (*q)->encryption_parameters.coeff_modulus == p->encryption_parameters.coeff_modulus,
          "PIRParameters wire round trip");
  }
  return 0;
}

static int string_encoder_tests() {
  const pir::EncryptionParameters ep = pir::GenerateEncryptionParams(4096);
  pir::StringEncoder enc(ep);
  // TestNumItemsPerPlaintext (string_encoder_test.cpp:64-71)
  CHECK(enc.num_items_per_plaintext(1) == 9728 && enc.num_items_per_plaintext(9728) == 1 &&
            enc.num_items_per_plaintext(9729) == 0 && enc.num_items_per_plaintext(99999) == 0 &&
            enc.num_items_per_plaintext(64) == 152 && enc.num_items_per_plaintext(288) == 33, "items per plaintext");
  {  // TestEncodeDecode (:73-83)
    const std::string value("This is a string test for random VALUES@!#");
    std::vector<uint64_t> pt;
    CHECK(enc.encode(value, pt).ok(), "encode");
    CHECK(pt.size() == (size_t)std::ceil((value.size() * 8) / 19.0), "coefficient count");
    for (uint64_t c : pt) CHECK(c < (1u << 19), "coefficient below 2^19");
    auto r = enc.decode(pt, pt.size() * 19 / 8);
    CHECK(r.ok() && r->size() >= value.size() && r->substr(0, value.size()) == value, "decode");
    for (size_t i = value.size(); i < r->size(); ++i) CHECK((*r)[i] == 0, "padding is zero");
  }
  std::mt19937_64 rng(42);
  {  // TestEncodeDecodePRN (:85-95): a full plaintext of random bytes
    std::string v(9728, 0);
    for (auto& ch : v) ch = (char)(rng() & 0xff);
    std::vector<uint64_t> pt;
    CHECK(enc.encode(v, pt).ok() && pt.size() == 4096, "full plaintext encode");
    auto r = enc.decode(pt, v.size());
    CHECK(r.ok() && *r == v, "full plaintext decode");
    std::string too_long(9729, 1);
    CHECK(!enc.encode(too_long, pt).ok(), "one byte too many must fail");
  }
  {  // TestEncodeDecodeVector (:97-118): 152 items of 64 bytes, decoded one by one at their offsets
    std::vector<std::string> v(152, std::string(64, 0));
    for (auto& s : v)
      for (auto& ch : s) ch = (char)(1 + rng() % 255);
    std::vector<uint64_t> pt;
    CHECK(enc.encode(v.begin(), v.end(), pt).ok(), "vector encode");
    size_t offset = 0;
    for (size_t i = 0; i < v.size(); ++i) {
      auto r = enc.decode(pt, v[i].size(), offset);
      offset += v[i].size();
      CHECK(r.ok() && *r == v[i], "vector decode");
    }
    CHECK(!enc.decode(pt, 64, 152 * 64).ok(), "decode beyond the data must fail");
  }
  return 0;
}

static int index_tests() {
  // CalculateIndicesTest (database_test.cpp:390-421): N=4096, 16-bit plain modulus
  struct Case { uint32_t n, size, d, index; U32 want; };
  const Case cases[] = {{100, 0, 1, 42, {42}},       {100, 0, 1, 7, {7}},        {84, 0, 2, 7, {0, 7}},
                        {87, 0, 2, 27, {3, 0}},      {87, 0, 2, 42, {4, 6}},     {87, 0, 2, 86, {9, 5}},
                        {82, 0, 3, 3, {0, 0, 3}},    {82, 0, 3, 20, {1, 0, 0}},  {82, 0, 3, 75, {3, 3, 3}},
                        {5000, 64, 1, 2222, {18}},   {5000, 64, 1, 1200, {10}}};
  for (const auto& c : cases) {
    auto p = pir::CreatePIRParameters(c.n, c.size, c.d, pir::GenerateEncryptionParams(4096, 16));
    CHECK(p.ok(), "CreatePIRParameters");
    CHECK(pir::PIRDatabase::calculate_indices(**p, c.index) == c.want, "calculate_indices");
  }
  // CalculateOffsetTest (:423-444)
  const uint32_t off[][4] = {{100, 0, 42, 0}, {1000, 64, 42, 2688}, {1000, 64, 960, 0}, {1000, 64, 999, 2496}};
  for (const auto& o : off) {
    auto p = pir::CreatePIRParameters(o[0], o[1], 1, pir::GenerateEncryptionParams(4096, 16));
    CHECK(p.ok() && pir::PIRDatabase::calculate_item_offset(**p, o[2]) == o[3], "calculate_item_offset");
  }
  // CalculateDimensionsTest (:446-464)
  struct Dim { uint32_t n, d; U32 want; };
  const Dim dims[] = {{100, 1, {100}},          {100, 2, {10, 10}},         {82, 2, {10, 9}},           {975, 2, {32, 31}},
                      {1000, 3, {10, 10, 10}},  {1001, 3, {11, 10, 10}},    {1000001, 3, {101, 100, 100}}};
  for (const auto& c : dims) CHECK(pir::PIRDatabase::calculate_dimensions(c.n, c.d) == c.want, "calculate_dimensions");
  // the BASELINE shapes
  CHECK(pir::PIRDatabase::calculate_dimensions(1639, 2) == (U32{41, 40}) &&
            pir::PIRDatabase::calculate_dimensions(110377, 2) == (U32{333, 332}) &&
            pir::PIRDatabase::calculate_dimensions(55189, 2) == (U32{235, 235}) &&
            pir::PIRDatabase::calculate_dimensions(441506, 2) == (U32{665, 664}), "BASELINE shapes");
  return 0;
}

// the fingerprint that finds a returning client's Galois keys again (PIRServer's key cache): deterministic, sensitive to
// every byte, to the length and to the salt that separates raw-limb keys from serialized ones
static int fingerprint_tests() {
  std::vector<uint64_t> a(5000), b;
  for (size_t i = 0; i < a.size(); ++i) a[i] = i * 0x9e3779b97f4a7c15ull + 12345;
  const auto f0 = pir::detail::fingerprint(a.data(), a.size() * 8);
  CHECK(f0 == pir::detail::fingerprint(a.data(), a.size() * 8), "deterministic");
  for (size_t pos : {size_t(0), size_t(1), size_t(31), size_t(32), size_t(4999 * 8 + 7), size_t(2500 * 8 + 3)}) {
    b = a;
    reinterpret_cast<unsigned char*>(b.data())[pos] ^= 1;
    CHECK(!(pir::detail::fingerprint(b.data(), b.size() * 8) == f0), "one flipped bit changes the fingerprint");
  }
  CHECK(!(pir::detail::fingerprint(a.data(), a.size() * 8 - 1) == f0), "length enters the fingerprint");
  CHECK(!(pir::detail::fingerprint(a.data(), a.size() * 8 - 8) == f0), "length enters the fingerprint (whole word)");
  CHECK(!(pir::detail::fingerprint(a.data(), a.size() * 8, 1) == f0), "salt enters the fingerprint");
  // tails shorter than a word and shorter than a 32-byte block
  for (size_t n : {size_t(0), size_t(1), size_t(7), size_t(8), size_t(9), size_t(33), size_t(63)}) {
    const auto f1 = pir::detail::fingerprint(a.data(), n);
    b = a;
    if (n) {
      reinterpret_cast<unsigned char*>(b.data())[n - 1] ^= 0x80;
      CHECK(!(pir::detail::fingerprint(b.data(), n) == f1), "last byte of a short input counts");
    }
    reinterpret_cast<unsigned char*>(b.data())[n] ^= 0x80;  // the byte after the end must not matter
    if (n) reinterpret_cast<unsigned char*>(b.data())[n - 1] ^= 0x80;
    CHECK(pir::detail::fingerprint(b.data(), n) == f1, "bytes beyond the length are not read");
  }
  return 0;
}

// serialization_test.cpp:80-164 restated on raw limbs: SaveRequest with and without keys, the objects survive the trip
static int serialization_tests() {
  const pir::EncryptionParameters ep = pir::GenerateEncryptionParams(4096);
  const size_t N = 4096, k = ep.coeff_modulus.size() - 1, key_limbs = k * 2 * (k + 1) * N;
  std::mt19937_64 rng(5);
  auto poly_limbs = [&](size_t polys, size_t n_mod, std::vector<uint64_t>& out) {
    out.resize(polys * n_mod * N);
    for (size_t p = 0; p < polys; ++p)
      for (size_t j = 0; j < n_mod; ++j)
        for (size_t n = 0; n < N; ++n) out[(p * n_mod + j) * N + n] = rng() % ep.coeff_modulus[j];
  };
  std::vector<std::vector<pir::Ciphertext>> cts(2, std::vector<pir::Ciphertext>(2));
  for (auto& q : cts)
    for (auto& ct : q) poly_limbs(2, k, ct.limbs);
  pir::GaloisKeys gk;
  gk.elts = pir::generate_galois_elts(N);
  gk.limbs.resize(gk.elts.size() * key_limbs);
  for (size_t e = 0; e < gk.elts.size() * k; ++e) {
    std::vector<uint64_t> digit;
    poly_limbs(2, k + 1, digit);
    std::copy(digit.begin(), digit.end(), gk.limbs.begin() + e * digit.size());
  }
  pir::RelinKeys rk;
  for (size_t j = 0; j < k; ++j) {
    std::vector<uint64_t> digit;
    poly_limbs(2, k + 1, digit);
    rk.limbs.insert(rk.limbs.end(), digit.begin(), digit.end());
  }
  // TestRequestSerialization (:135-164): all three fields
  pir::wire::RequestMsg m;
  CHECK(pir::SaveRequest(ep, cts, gk, rk, &m).ok(), "SaveRequest");
  pir::wire::RequestMsg back;
  CHECK(pir::wire::Parse(pir::wire::Serialize(m), &back) && back.query.size() == 2 && back.query[1].ct.size() == 2,
        "request parse");
  for (size_t q = 0; q < 2; ++q) {
    auto loaded = pir::LoadCiphertexts(ep, back.query[q]);
    CHECK(loaded.ok() && loaded->size() == 2 && (*loaded)[0].limbs == cts[q][0].limbs && (*loaded)[1].limbs == cts[q][1].limbs,
          "ciphertexts round trip");
  }
  auto gk2 = pir::DeserializeGaloisKeys(ep, back.galois_keys);
  CHECK(gk2.ok() && gk2->elts.size() == gk.elts.size(), "galois keys round trip: has_key for every element");
  for (size_t e = 0; e < gk.elts.size(); ++e) {  // returned in slot order
    size_t at = 0;
    while (at < gk2->elts.size() && gk2->elts[at] != gk.elts[e]) ++at;
    CHECK(at < gk2->elts.size() && std::equal(gk.limbs.begin() + e * key_limbs, gk.limbs.begin() + (e + 1) * key_limbs,
                                               gk2->limbs.begin() + at * key_limbs), "galois key limbs");
  }
  auto rk2 = pir::DeserializeRelinKeys(ep, back.relin_keys);
  CHECK(rk2.ok() && rk2->limbs == rk.limbs, "relin keys round trip");
  CHECK(!pir::DeserializeRelinKeys(ep, back.galois_keys).ok(), "a Galois key set is not a relinearization key");
  CHECK(!pir::DeserializeRelinKeys(ep, back.relin_keys.substr(0, back.relin_keys.size() - 3)).ok(), "truncated relin keys");
  // TestRequestSerialization_Shortcut (:115-132): no keys at all
  pir::wire::RequestMsg sc;
  CHECK(pir::SaveRequest(ep, cts, &sc).ok() && sc.galois_keys.empty() && sc.relin_keys.empty() && sc.query.size() == 2,
        "SaveRequest shortcut");
  CHECK(!pir::SaveRequest(ep, cts, nullptr).ok() && !pir::SaveCiphertexts(ep, cts[0], nullptr).ok(), "output nullptr");
  // a size-3 ciphertext (ciphertext-multiplication reply without relinearization) serializes with its size
  pir::Ciphertext three;
  poly_limbs(3, k, three.limbs);
  pir::wire::CiphertextData d;
  std::string err;
  CHECK(pir::wire::LoadCiphertext(pir::SerializeCiphertext(ep, three), (uint32_t)N, ep.coeff_modulus.data(), k, &d, &err) &&
            d.size == 3 && d.limbs == three.limbs, "size-3 ciphertext round trip");
  return 0;
}

int main() {
  if (parameters_tests() || string_encoder_tests() || index_tests() || fingerprint_tests() || serialization_tests()) return 1;
  // without a device the factories must fail loudly, never fall back (server.cpp:35-42 shape)
  auto p = *pir::CreatePIRParameters(10, 0, 1);
  auto db = pir::PIRDatabase::Create(p);
  if (!db.ok()) std::printf("no CUDA device here: PIRDatabase::Create -> status %d (%s)\n", db.status().code(),
                            db.status().message().c_str());
  std::printf("SHIM_HOST_TEST_OK\n");
  return 0;
}
