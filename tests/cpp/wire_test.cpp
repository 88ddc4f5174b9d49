// CPU-only test of the wire layer (pir_b200/cpp/wire.hpp) end to end against the oracle: a request as the reference
// client would send it — query ciphertext + SEED-COMPRESSED Galois keys + relinearization keys, protobuf-framed — is
// parsed back into raw limbs, answered by the CPU oracle (the checker; no GPU involved), the reply is serialized,
// parsed again and decrypted.  Pins, without a device: framing, SEAL object save/load, seed expansion producing keys
// that actually key-switch, reply serialization.  (server_test.cpp:98-121 shape: 10 items, d=1.)
#include <cstdio>
#include <random>

#include "../../oracle/pir_oracle.hpp"
#include "../../pir_b200/cpp/wire.hpp"

#define CHECK(cond, msg)                                                \
  do {                                                                  \
    if (!(cond)) {                                                      \
      std::fprintf(stderr, "FAIL %s:%d %s\n", __FILE__, __LINE__, msg); \
      return 1;                                                         \
    }                                                                   \
  } while (0)

namespace w = pir::wire;

int main() {
  const uint32_t N = 4096;
  const std::vector<uint64_t> mods = orc::bfv_default_coeff_modulus(N);
  const uint64_t t = orc::plain_modulus_batching(N, 20);
  orc::Context octx(N, mods, t);
  orc::Crypto crypto(octx);
  orc::Rng rng(123);
  std::mt19937_64 seeder(99);
  const size_t k = octx.k, K1 = k + 1, per_digit = 2 * K1 * N;
  orc::SecretKey sk = orc::gen_secret_key(octx, rng);
  orc::PublicKey pk = orc::gen_public_key(octx, sk, rng);

  w::SealParams sp;
  sp.poly_modulus_degree = N;
  sp.coeff_modulus = mods;
  sp.plain_modulus = t;

  // ---- client: Galois keys, seed-compressed (c1 := expansion of a fresh seed, c0 fixed up accordingly) ----
  const std::vector<uint32_t> elts = orc::generate_galois_elts(N);
  w::KSwitchKeysData K;
  K.parms_id = w::key_parms_id(sp);
  K.keys.assign(N, {});
  std::vector<std::vector<w::seed_type>> seeds(N);
  std::vector<uint64_t> key(k * per_digit), a2(K1 * N);
  for (uint32_t g : elts) {
    orc::gen_galois_key(octx, sk, g, rng, key.data());
    const uint32_t idx = w::galois_index(g);
    K.keys[idx].resize(k);
    seeds[idx].resize(k);
    for (size_t J = 0; J < k; ++J) {
      for (auto& x : seeds[idx][J]) x = seeder();
      w::BlakePRNG prng(seeds[idx][J]);
      w::sample_poly_uniform(prng, N, mods.data(), K1, a2.data());
      uint64_t* kj = key.data() + J * per_digit;
      for (size_t j = 0; j < K1; ++j) {
        const orc::Modulus& m = octx.mod(j);
        for (size_t n = 0; n < N; ++n) {
          uint64_t& c0 = kj[j * N + n];
          uint64_t& c1 = kj[(K1 + j) * N + n];
          c0 = orc::addmod(c0, orc::mulmod(orc::submod(c1, a2[j * N + n], m), sk.ntt[j * N + n], m), m);
          c1 = a2[j * N + n];
        }
      }
      auto& ct = K.keys[idx][J];
      ct.parms_id = K.parms_id;
      ct.is_ntt_form = true;
      ct.size = 2;
      ct.poly_modulus_degree = N;
      ct.coeff_modulus_size = K1;
      ct.limbs.assign(kj, kj + per_digit);
    }
  }
  // ---- client: query for index 7 of 10 items (client.cpp:92-144: value m^-1 at the index) ----
  const size_t n_items = 10, index = 7;
  const uint64_t m_pow = orc::next_power_two<uint64_t>(n_items);
  std::vector<uint64_t> pt(N, 0);
  pt[index] = orc::invmod(m_pow % t, orc::Modulus(t));
  w::CiphertextData qct;
  qct.parms_id = w::data_parms_id(sp);
  qct.size = 2;
  qct.poly_modulus_degree = N;
  qct.coeff_modulus_size = k;
  qct.limbs.resize(octx.ct_limbs());
  crypto.encrypt(pk, pt.data(), N, rng, qct.limbs.data());

  w::RequestMsg req;
  req.query.emplace_back();
  req.query.back().ct.push_back(w::SaveCiphertext(qct));
  req.galois_keys = w::SaveKSwitchKeys(K, &seeds);
  {
    w::KSwitchKeysData R;  // relinearization keys: one slot, also seed-compressed
    R.parms_id = K.parms_id;
    R.keys.push_back(K.keys[w::galois_index(elts[0])]);
    std::vector<std::vector<w::seed_type>> rs{seeds[w::galois_index(elts[0])]};
    req.relin_keys = w::SaveKSwitchKeys(R, &rs);
  }
  const std::string wire_request = w::Serialize(req);
  std::printf("request: %zu bytes (galois keys %zu, relin keys %zu, query %zu)\n", wire_request.size(),
              req.galois_keys.size(), req.relin_keys.size(), req.query[0].ct[0].size());

  // ---- server side: parse everything back into raw limbs ----
  w::RequestMsg got;
  CHECK(w::Parse(wire_request, &got), "Request must parse");
  CHECK(got.query.size() == 1 && got.query[0].ct.size() == 1, "one query, one ciphertext");
  std::string err;
  w::KSwitchKeysData GK, RK;
  CHECK(w::LoadKSwitchKeys(got.galois_keys, sp, &GK, &err), err.c_str());
  CHECK(w::LoadKSwitchKeys(got.relin_keys, sp, &RK, &err, false), err.c_str());
  CHECK(GK.parms_id == w::key_parms_id(sp), "key parms_id");
  orc::GaloisKeys ogk;
  for (size_t s = 0; s < GK.keys.size(); ++s) {
    if (GK.keys[s].empty()) continue;
    ogk.elts.push_back(w::galois_elt_of_index((uint32_t)s));
    for (const auto& ct : GK.keys[s]) {
      CHECK(ct.was_seeded, "keys must have travelled seed-compressed");
      ogk.data.insert(ogk.data.end(), ct.limbs.begin(), ct.limbs.end());
    }
  }
  CHECK(ogk.elts.size() == elts.size(), "every Galois element present");
  w::CiphertextData q2;
  CHECK(w::LoadCiphertext(got.query[0].ct[0], N, mods.data(), k, &q2, &err), err.c_str());
  CHECK(q2.limbs == qct.limbs && q2.parms_id == qct.parms_id && !q2.is_ntt_form, "query ciphertext round trip");

  // ---- the checker answers (database of integers 0, 10, 20, ...: IntegerEncoder plaintexts) ----
  std::vector<uint64_t> db(n_items * octx.pt_limbs());
  for (size_t i = 0; i < n_items; ++i) {
    std::vector<uint64_t> c(N, 0);
    for (uint64_t v = 10 * i, b = 0; v; v >>= 1, ++b)
      if (v & 1) c[b] = 1;
    orc::plain_to_ntt(octx, c.data(), N, db.data() + i * octx.pt_limbs());
  }
  const uint32_t dims[1] = {(uint32_t)n_items};
  std::vector<uint64_t> reply;
  CHECK(orc::process_query(octx, db.data(), n_items, dims, 1, ogk, q2.limbs.data(), 1, reply) == 0, "process_query");
  CHECK(reply.size() == octx.ct_limbs(), "one reply ciphertext");

  // ---- reply over the wire and back ----
  w::ResponseMsg resp;
  resp.reply.emplace_back();
  w::CiphertextData rct;
  rct.parms_id = q2.parms_id;
  rct.size = 2;
  rct.poly_modulus_degree = N;
  rct.coeff_modulus_size = k;
  rct.limbs = reply;
  resp.reply.back().ct.push_back(w::SaveCiphertext(rct));
  const std::string wire_response = w::Serialize(resp);
  w::ResponseMsg resp2;
  CHECK(w::Parse(wire_response, &resp2) && resp2.reply.size() == 1 && resp2.reply[0].ct.size() == 1, "Response parse");
  w::CiphertextData r2;
  CHECK(w::LoadCiphertext(resp2.reply[0].ct[0], N, mods.data(), k, &r2, &err), err.c_str());
  std::vector<uint64_t> dec(N);
  crypto.decrypt(sk, r2.limbs.data(), dec.data());
  // IntegerEncoder::decode: sum of coeff * 2^i (coefficients here are 0/1 after the m^-1 scaling cancels)
  uint64_t value = 0;
  for (size_t i = 0; i < 63; ++i) value += dec[i] << i;
  CHECK(value == 10 * index, "decrypted reply must be the selected database entry");

  // ---- the copy-free variants the device server uses must be byte-for-byte the object-based ones ----
  {
    w::RequestView view;
    CHECK(w::ParseView(wire_request, &view), "ParseView");
    CHECK(view.query.size() == got.query.size() && view.query[0].size() == got.query[0].ct.size(), "view: query shape");
    CHECK(view.query[0][0] == std::string_view(got.query[0].ct[0]), "view: ciphertext bytes");
    CHECK(view.galois_keys == std::string_view(got.galois_keys), "view: galois keys");
    CHECK(view.relin_keys == std::string_view(got.relin_keys), "view: relin keys");
    CHECK(view.galois_keys.data() >= wire_request.data() &&
              view.galois_keys.data() + view.galois_keys.size() <= wire_request.data() + wire_request.size(),
          "view must point into the request buffer");
    w::RequestView bad;
    CHECK(!w::ParseView(std::string_view(wire_request).substr(0, wire_request.size() - 5), &bad),
          "truncated request must not parse as a view");

    // LoadCiphertextTo: limbs land in the caller's buffer, metadata in `meta`, same validation
    std::vector<uint64_t> direct(octx.ct_limbs(), ~0ull);
    w::CiphertextData meta;
    CHECK(w::LoadCiphertextTo(view.query[0][0], N, mods.data(), k, direct.data(), &meta, &err), err.c_str());
    CHECK(direct == qct.limbs && meta.limbs.empty() && meta.parms_id == qct.parms_id && !meta.is_ntt_form && meta.size == 2,
          "LoadCiphertextTo: limbs / metadata");
    // a seed-compressed query ciphertext expands into the same buffer exactly as LoadCiphertext expands it
    w::seed_type s{};
    for (auto& x : s) x = seeder();
    const std::string seeded_blob = w::SaveCiphertext(qct, &s);
    w::CiphertextData via_obj;
    CHECK(w::LoadCiphertext(seeded_blob, N, mods.data(), k, &via_obj, &err) && via_obj.was_seeded, err.c_str());
    std::fill(direct.begin(), direct.end(), ~0ull);
    CHECK(w::LoadCiphertextTo(seeded_blob, N, mods.data(), k, direct.data(), &meta, &err) && meta.was_seeded, err.c_str());
    CHECK(direct == via_obj.limbs, "seeded ciphertext: in-place expansion differs");
    // wrong size (3 polynomials) and out-of-range limbs are refused, as by the object path
    w::CiphertextData three = qct;
    three.size = 3;
    three.limbs.resize(3 * k * N, 1);
    CHECK(!w::LoadCiphertextTo(w::SaveCiphertext(three), N, mods.data(), k, direct.data(), &meta, &err),
          "a 3-polynomial object must not load into a 2-polynomial slot");
    w::CiphertextData big = qct;
    big.limbs[5] = mods[0];
    CHECK(!w::LoadCiphertextTo(w::SaveCiphertext(big), N, mods.data(), k, direct.data(), &meta, &err),
          "limb >= modulus must be refused");

    // AppendCiphertextBlob == SaveCiphertext, and the hand-framed response == Serialize(ResponseMsg)
    std::string blob;
    w::AppendCiphertextBlob(blob, reply.data(), 2, N, k, rct.parms_id, false);
    const std::string want = w::SaveCiphertext(rct);
    CHECK(blob == want && blob.size() == w::CiphertextBlobSize(2, N, k), "AppendCiphertextBlob bytes");
    // Response{ reply: [ Ciphertexts{ ct: [blob, blob] }, Ciphertexts{ ct: [blob] } ] } framed by hand
    w::ResponseMsg multi;
    multi.reply.resize(2);
    multi.reply[0].ct = {want, want};
    multi.reply[1].ct = {want};
    std::string framed;
    const size_t bsz = w::CiphertextBlobSize(2, N, k);
    for (size_t n_ct : {size_t(2), size_t(1)}) {
      const size_t inner = n_ct * (1 + w::varint_size(bsz) + bsz);
      w::put_varint(framed, (1u << 3) | 2);
      w::put_varint(framed, inner);
      for (size_t c = 0; c < n_ct; ++c) {
        w::put_varint(framed, (1u << 3) | 2);
        w::put_varint(framed, bsz);
        w::AppendCiphertextBlob(framed, reply.data(), 2, N, k, rct.parms_id, false);
      }
    }
    CHECK(framed == w::Serialize(multi), "hand-framed response differs from Serialize(ResponseMsg)");
  }
  // ---- the reference's own serialization tests, restated (pir/cpp/serialization_test.cpp) ----
  {
    // TestResponseSerialization (:62-78): encode 987654321, encrypt, save into a Response, reload, decrypt
    std::vector<uint64_t> ipt(N, 0), back(N);
    for (uint64_t v = 987654321ull, b = 0; v; v >>= 1, ++b)
      if (v & 1) ipt[b] = 1;
    w::CiphertextData c1;
    c1.parms_id = w::data_parms_id(sp);
    c1.size = 2;
    c1.poly_modulus_degree = N;
    c1.coeff_modulus_size = k;
    c1.limbs.resize(octx.ct_limbs());
    crypto.encrypt(pk, ipt.data(), N, rng, c1.limbs.data());
    w::ResponseMsg rp;
    rp.reply.emplace_back();
    rp.reply.back().ct.push_back(w::SaveCiphertext(c1));
    w::ResponseMsg rp2;
    CHECK(w::Parse(w::Serialize(rp), &rp2) && rp2.reply.size() == 1 && rp2.reply[0].ct.size() == 1, "reload size");
    w::CiphertextData c2;
    CHECK(w::LoadCiphertext(rp2.reply[0].ct[0], N, mods.data(), k, &c2, &err), err.c_str());
    crypto.decrypt(sk, c2.limbs.data(), back.data());
    CHECK(back == ipt, "TestResponseSerialization: reloaded plaintext differs");

    // TestRequestSerialization_IndividualMethods (:80-113): galois_keys_local (not seed-compressed) + relin keys;
    // after the round trip every Galois element must be present (has_key)
    w::KSwitchKeysData L;
    L.parms_id = w::key_parms_id(sp);
    L.keys.assign(N, {});
    for (uint32_t g : elts) L.keys[w::galois_index(g)] = K.keys[w::galois_index(g)];
    w::RequestMsg rq;
    rq.query.emplace_back();
    rq.query.back().ct.push_back(w::SaveCiphertext(c1));
    rq.galois_keys = w::SaveKSwitchKeys(L);  // full ciphertexts, no seeds
    w::KSwitchKeysData RL;
    RL.parms_id = L.parms_id;
    RL.keys.push_back(K.keys[w::galois_index(elts[1])]);
    rq.relin_keys = w::SaveKSwitchKeys(RL);
    w::RequestMsg rq2;
    CHECK(w::Parse(w::Serialize(rq), &rq2), "request parse");
    w::KSwitchKeysData L2, RL2;
    CHECK(w::LoadKSwitchKeys(rq2.galois_keys, sp, &L2, &err), err.c_str());
    for (uint32_t g : elts) {
      const uint32_t idx = w::galois_index(g);
      CHECK(idx < L2.keys.size() && L2.keys[idx].size() == k, "has_key");
      for (size_t J = 0; J < k; ++J)
        CHECK(!L2.keys[idx][J].was_seeded && L2.keys[idx][J].limbs == K.keys[idx][J].limbs, "key limbs round trip");
    }
    CHECK(w::LoadKSwitchKeys(rq2.relin_keys, sp, &RL2, &err) && RL2.keys.size() == 1, "relin keys reload");
    CHECK(rq.galois_keys.size() > 18 * req.galois_keys.size() / 10, "seed compression must halve the key bytes");

    // TestRequestSerialization_Shortcut (:115-132): SaveRequest(cts) leaves both key fields empty
    w::RequestMsg sc;
    sc.query.emplace_back();
    sc.query.back().ct.push_back(w::SaveCiphertext(c1));
    w::RequestMsg sc2;
    CHECK(w::Parse(w::Serialize(sc), &sc2), "shortcut parse");
    CHECK(sc2.query.size() == 1 && sc2.galois_keys.empty() && sc2.relin_keys.empty(), "shortcut: no keys");
    // a server must refuse it: there is nothing to expand with (server.cpp:46-48 fails to deserialize "")
    w::KSwitchKeysData none;
    CHECK(!w::LoadKSwitchKeys(sc2.galois_keys, sp, &none, &err), "empty galois_keys must not load");
  }
  std::printf("WIRE_TEST_OK: seeded keys expanded, query answered by the oracle, reply decrypts to %llu\n",
              (unsigned long long)value);
  return 0;
}
