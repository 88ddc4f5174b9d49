// End-to-end test of the C++ shim (pir_b200/cpp/pir_b200.hpp) written like the reference's own tests
// (pir/cpp/correctness_test.cpp:82-93, pir/cpp/server_test.cpp:98-121, 209-260): client -> PIRServer::ProcessRequest
// -> client.  The client side (keygen / encrypt / decrypt) and the expected answers come from the CPU oracle
// (oracle/pir_oracle.hpp), which this test is allowed to use as the checker.
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../../oracle/pir_oracle.hpp"
#include "../../pir_b200/cpp/pir_b200.hpp"

#define CHECK(cond, msg)                                            \
  do {                                                              \
    if (!(cond)) {                                                  \
      std::fprintf(stderr, "FAIL %s:%d %s\n", __FILE__, __LINE__, msg); \
      return 1;                                                     \
    }                                                               \
  } while (0)

static int run_case(size_t dbsize, size_t elem_size, size_t d, size_t desired_index, size_t shards = 1) {
  auto ep = pir::GenerateEncryptionParams(4096, 20);
  auto params_or = pir::CreatePIRParameters(dbsize, elem_size, d, ep);
  CHECK(params_or.ok(), "CreatePIRParameters");
  auto params = *params_or;
  std::mt19937_64 rng(42);
  std::vector<std::string> db(dbsize, std::string(params->bytes_per_item, 0));
  for (auto& s : db)
    for (auto& ch : s) ch = (char)(rng() & 0xff);
  // shards > 1: the database is split by rows over `shards` contexts of this process (all on GPU 0 here; one per GPU
  // in a deployment) and the server runs the peer-memory exchange flow (pirb_dist_*) behind the same ProcessRequest
  auto db_or = pir::PIRDatabase::Create(db, params, std::vector<int>(shards, 0));
  CHECK(db_or.ok(), db_or.status().message().c_str());
  CHECK((*db_or)->shard_count() == shards, "shard count");
  auto server_or = pir::PIRServer::Create(*db_or, params, /*max_queries_per_gpu=*/2);
  CHECK(server_or.ok(), server_or.status().message().c_str());
  auto& server = *server_or;

  // ---- client side (oracle harness) ----
  orc::Context octx(ep.poly_modulus_degree, ep.coeff_modulus, ep.plain_modulus);
  orc::Crypto crypto(octx);
  orc::Rng r1(7);
  orc::SecretKey sk = orc::gen_secret_key(octx, r1);
  orc::PublicKey pk = orc::gen_public_key(octx, sk, r1);
  pir::GaloisKeys gk;
  gk.elts = pir::generate_galois_elts(ep.poly_modulus_degree);
  const size_t key_limbs = octx.k * 2 * (octx.k + 1) * octx.N;
  gk.limbs.resize(gk.elts.size() * key_limbs);
  for (size_t i = 0; i < gk.elts.size(); ++i) orc::gen_galois_key(octx, sk, gk.elts[i], r1, gk.limbs.data() + i * key_limbs);

  // query packing as PIRClient::createQueryFor (client.cpp:92-144) for dim_sum <= N
  const size_t N = octx.N;
  auto indices = (*db_or)->calculate_indices((uint32_t)desired_index);
  size_t dim_sum = 0;
  for (auto v : params->dimensions) dim_sum += v;
  CHECK(dim_sum < N, "test shape must fit one query ciphertext");
  const uint64_t m = pirb_next_power_two(dim_sum);
  const uint64_t m_inv = orc::invmod(m % ep.plain_modulus, orc::Modulus(ep.plain_modulus));
  std::vector<uint64_t> pt(N, 0);
  size_t offset = 0;
  for (size_t i = 0; i < indices.size(); ++i) {
    pt[offset + indices[i]] = m_inv;
    offset += params->dimensions[i];
  }
  pir::Request req;
  req.galois_keys = gk;
  req.query.resize(1);
  req.query[0].resize(1);
  req.query[0][0].limbs.resize(octx.ct_limbs());
  crypto.encrypt(pk, pt.data(), N, r1, req.query[0][0].data());

  auto resp_or = server->ProcessRequest(req);
  CHECK(resp_or.ok(), resp_or.status().message().c_str());
  auto& reply = resp_or->reply[0];

  // ---- oracle answer on the same inputs: limbs must be identical ----
  std::vector<uint64_t> coeffs, db_ntt(params->num_pt * octx.pt_limbs());
  pir::StringEncoder enc(ep);
  for (size_t i = 0; i < params->num_pt; ++i) {
    auto b = db.begin() + i * params->items_per_plaintext;
    auto e = (size_t)(db.end() - b) > params->items_per_plaintext ? b + params->items_per_plaintext : db.end();
    CHECK(enc.encode(b, e, coeffs).ok(), "encode");
    orc::plain_to_ntt(octx, coeffs.data(), coeffs.size(), db_ntt.data() + i * octx.pt_limbs());
  }
  orc::GaloisKeys ogk;
  ogk.elts = gk.elts;
  ogk.data = gk.limbs;
  std::vector<uint64_t> want;
  int rc = orc::process_query(octx, db_ntt.data(), params->num_pt, params->dimensions.data(), params->dimensions.size(),
                              ogk, req.query[0][0].data(), 1, want);
  CHECK(rc == 0, "oracle process_query");
  CHECK(want.size() == reply.size() * octx.ct_limbs(), "reply count");
  for (size_t i = 0; i < reply.size(); ++i)
    CHECK(std::memcmp(reply[i].data(), want.data() + i * octx.ct_limbs(), octx.ct_limbs() * 8) == 0,
          "GPU reply differs from the oracle");

  // ---- a batch of five copies of the query (more than ranks x queries per GPU when sharded: two steps, the last one
  // padded), keys found again in the server's cache ----
  {
    pir::Request req5 = req;
    req5.query.assign(5, req.query[0]);
    const size_t hits_before = server->key_cache_hits();
    auto r5 = server->ProcessRequest(req5);
    CHECK(r5.ok(), r5.status().message().c_str());
    CHECK(server->key_cache_hits() == hits_before + 1, "Galois keys of a returning client must come from the cache");
    CHECK(r5->reply.size() == 5, "one reply per query");
    for (size_t qi = 0; qi < 5; ++qi) {
      CHECK(r5->reply[qi].size() == reply.size(), "reply count in batch");
      for (size_t i = 0; i < reply.size(); ++i)
        CHECK(std::memcmp(r5->reply[qi][i].data(), want.data() + i * octx.ct_limbs(), octx.ct_limbs() * 8) == 0,
              "batched reply differs from the oracle");
    }
  }

  // ---- the same request over the reference's wire format (serialization.h:81-138, payload.proto) ----
  // Keys travel seed-compressed as the reference client sends them (client.cpp:47-54): replace every key's uniform
  // polynomial a by the expansion a' of a fresh seed and fix c0 so that it is still a valid key:
  // c0' = c0 + (a - a') * s (NTT domain).  The server must rebuild a' from the seed alone.
  {
    const size_t K1 = octx.k + 1, per_digit = 2 * K1 * N;
    const pir::wire::SealParams sp = pir::ToSealParams(ep);
    pir::wire::KSwitchKeysData K;
    K.parms_id = pir::wire::key_parms_id(sp);
    K.keys.assign(N, {});
    std::vector<std::vector<pir::wire::seed_type>> seeds(N);
    pir::GaloisKeys gk2 = gk;
    std::vector<uint64_t> a2(K1 * N);
    for (size_t e = 0; e < gk.elts.size(); ++e) {
      const uint32_t idx = pir::wire::galois_index(gk.elts[e]);
      K.keys[idx].resize(octx.k);
      seeds[idx].resize(octx.k);
      for (size_t J = 0; J < octx.k; ++J) {
        for (auto& w : seeds[idx][J]) w = rng();
        pir::wire::BlakePRNG prng(seeds[idx][J]);
        pir::wire::sample_poly_uniform(prng, (uint32_t)N, ep.coeff_modulus.data(), K1, a2.data());
        uint64_t* kj = gk2.limbs.data() + (e * octx.k + J) * per_digit;
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
    pir::wire::RequestMsg msg;
    msg.galois_keys = pir::wire::SaveKSwitchKeys(K, &seeds);
    CHECK(msg.galois_keys.size() < pir::SerializeGaloisKeys(ep, gk2).size() * 6 / 10, "seeded keys must be ~half the size");
    {  // relin keys are always sent (client.cpp:49,53-54) and only parsed: one slot of k seeded keys
      pir::wire::KSwitchKeysData R;
      R.parms_id = K.parms_id;
      R.keys.push_back(K.keys[pir::wire::galois_index(gk.elts[0])]);
      std::vector<std::vector<pir::wire::seed_type>> rs{seeds[pir::wire::galois_index(gk.elts[0])]};
      msg.relin_keys = pir::wire::SaveKSwitchKeys(R, &rs);
    }
    msg.query.emplace_back();
    msg.query.back().ct.push_back(pir::SerializeCiphertext(ep, req.query[0][0]));
    const std::string wire_req = pir::wire::Serialize(msg);
    auto wire_resp = server->ProcessRequest(wire_req);
    CHECK(wire_resp.ok(), wire_resp.status().message().c_str());
    pir::wire::ResponseMsg rmsg;
    CHECK(pir::wire::Parse(*wire_resp, &rmsg) && rmsg.reply.size() == 1, "Response parse");
    orc::GaloisKeys ogk2;
    ogk2.elts = gk2.elts;
    ogk2.data = gk2.limbs;
    std::vector<uint64_t> want2;
    CHECK(orc::process_query(octx, db_ntt.data(), params->num_pt, params->dimensions.data(), params->dimensions.size(),
                             ogk2, req.query[0][0].data(), 1, want2) == 0, "oracle process_query (seeded keys)");
    CHECK(rmsg.reply[0].ct.size() * octx.ct_limbs() == want2.size(), "wire reply count");
    for (size_t i = 0; i < rmsg.reply[0].ct.size(); ++i) {
      pir::wire::parms_id_type pid;
      auto ct = pir::DeserializeCiphertext(ep, rmsg.reply[0].ct[i], &pid);
      CHECK(ct.ok(), "reply ciphertext must deserialize");
      CHECK(pid == pir::wire::data_parms_id(sp), "reply parms_id is the query's");
      CHECK(std::memcmp(ct->data(), want2.data() + i * octx.ct_limbs(), octx.ct_limbs() * 8) == 0,
            "wire-path reply differs from the oracle");
    }
    // malformed requests are InvalidArgument, as SEALDeserialize failures are (serialization.h:113-115)
    auto bad1 = server->ProcessRequest(wire_req.substr(0, wire_req.size() / 2));
    CHECK(!bad1.ok() && bad1.status().code() == PIRB_INVALID_ARGUMENT, "truncated request");
    pir::wire::RequestMsg m2 = msg;
    m2.query[0].ct[0][100] ^= 0x40;  // high byte of a limb -> limb >= modulus
    m2.query[0].ct[0][111] = (char)0xff;
    auto bad2 = server->ProcessRequest(pir::wire::Serialize(m2));
    CHECK(!bad2.ok() && bad2.status().code() == PIRB_INVALID_ARGUMENT, "out-of-range ciphertext limb");
    pir::wire::RequestMsg m3 = msg;
    m3.galois_keys.clear();
    auto bad3 = server->ProcessRequest(pir::wire::Serialize(m3));
    CHECK(!bad3.ok() && bad3.status().code() == PIRB_INVALID_ARGUMENT, "missing galois keys");
    // PIRParameters over the wire (parameters.cpp:100-101)
    auto p2 = pir::ParsePIRParameters(pir::SerializePIRParameters(*params));
    CHECK(p2.ok() && (*p2)->num_pt == params->num_pt && (*p2)->dimensions == params->dimensions &&
              (*p2)->encryption_parameters.coeff_modulus == ep.coeff_modulus &&
              (*p2)->encryption_parameters.plain_modulus == ep.plain_modulus &&
              (*p2)->bytes_per_item == params->bytes_per_item, "PIRParameters wire round trip");
  }

  // ---- client decode (client.cpp:219-255) ----
  const size_t two_er = 2 * octx.expansion_ratio();
  std::vector<std::vector<uint64_t>> cts;
  for (auto& c : reply) cts.push_back(c.limbs);
  std::vector<std::vector<uint64_t>> pts;
  for (size_t level = 0; level < d; ++level) {
    pts.assign(cts.size(), std::vector<uint64_t>(N));
    for (size_t i = 0; i < cts.size(); ++i) crypto.decrypt(sk, cts[i].data(), pts[i].data());
    if (pts.size() <= 1) break;
    std::vector<std::vector<uint64_t>> next(cts.size() / two_er, std::vector<uint64_t>(octx.ct_limbs(), 0));
    for (size_t i = 0; i < next.size(); ++i) {
      std::vector<uint64_t> flat(two_er * N);
      for (size_t e = 0; e < two_er; ++e) std::copy(pts[i * two_er + e].begin(), pts[i * two_er + e].end(), flat.begin() + e * N);
      orc::reencode_decode(octx, flat.data(), next[i].data());
    }
    cts.swap(next);
  }
  auto got = enc.decode(pts[0], params->bytes_per_item, (*db_or)->calculate_item_offset((uint32_t)desired_index));
  CHECK(got.ok(), "decode");
  CHECK(*got == db[desired_index], "retrieved element differs");
  std::printf("ok: %zu items x %u B, d=%zu, index %zu, %zu shard(s): reply (%zu cts) bit-exact vs oracle, element recovered\n",
              dbsize, params->bytes_per_item, d, desired_index, shards, reply.size());
  return 0;
}

// correctness_test.cpp:94-105 / server_test.cpp:209-260 with use_ciphertext_multiplication: relinearization keys in the
// request, ONE reply ciphertext of size 2 per query (client.cpp:196-217), bit-exact against the oracle's restatement of
// Evaluator::multiply + relinearize_inplace; the serialized form of the same request; a request without the keys.
static int run_ct_case(size_t dbsize, size_t d, size_t desired_index, int plain_bits, size_t bits_per_coeff) {
  auto ep = pir::GenerateEncryptionParams(4096, plain_bits);
  auto params_or = pir::CreatePIRParameters(dbsize, 0, d, ep, /*use_ciphertext_multiplication=*/true, bits_per_coeff);
  CHECK(params_or.ok() && (*params_or)->use_ciphertext_multiplication, "CreatePIRParameters (CT mode)");
  auto params = *params_or;
  std::mt19937_64 rng(43);
  std::vector<std::string> db(dbsize, std::string(params->bytes_per_item, 0));
  for (auto& s : db)
    for (auto& ch : s) ch = (char)(rng() & 0xff);
  auto db_or = pir::PIRDatabase::Create(db, params);
  CHECK(db_or.ok(), db_or.status().message().c_str());
  auto server_or = pir::PIRServer::Create(*db_or, params);
  CHECK(server_or.ok(), server_or.status().message().c_str());
  auto& server = *server_or;
  {  // this mode runs on one GPU
    auto two = pir::PIRDatabase::Create(params, std::vector<int>{0, 0});
    CHECK(!two.ok() && two.status().code() == PIRB_INVALID_ARGUMENT, "sharded CT-mode database must be refused");
  }
  orc::Context octx(ep.poly_modulus_degree, ep.coeff_modulus, ep.plain_modulus);
  orc::Crypto crypto(octx);
  orc::RnsTool rns(octx);
  orc::Rng r1(9);
  orc::SecretKey sk = orc::gen_secret_key(octx, r1);
  orc::PublicKey pk = orc::gen_public_key(octx, sk, r1);
  const size_t N = octx.N, key_limbs = octx.k * 2 * (octx.k + 1) * N;
  pir::GaloisKeys gk;
  gk.elts = pir::generate_galois_elts(N);
  gk.limbs.resize(gk.elts.size() * key_limbs);
  for (size_t i = 0; i < gk.elts.size(); ++i) orc::gen_galois_key(octx, sk, gk.elts[i], r1, gk.limbs.data() + i * key_limbs);
  pir::RelinKeys relin;
  relin.limbs.resize(key_limbs);
  orc::gen_relin_key(octx, sk, r1, relin.limbs.data());

  auto indices = (*db_or)->calculate_indices((uint32_t)desired_index);
  size_t dim_sum = 0;
  for (auto v : params->dimensions) dim_sum += v;
  const uint64_t m_inv = orc::invmod(pirb_next_power_two(dim_sum) % ep.plain_modulus, orc::Modulus(ep.plain_modulus));
  std::vector<uint64_t> pt(N, 0);
  size_t offset = 0;
  for (size_t i = 0; i < indices.size(); ++i) {
    pt[offset + indices[i]] = m_inv;
    offset += params->dimensions[i];
  }
  pir::Request req;
  req.galois_keys = gk;
  req.relin_keys = relin;
  req.query.assign(2, std::vector<pir::Ciphertext>(1));  // the same index twice, two fresh encryptions
  for (auto& q : req.query) {
    q[0].limbs.resize(octx.ct_limbs());
    crypto.encrypt(pk, pt.data(), N, r1, q[0].data());
  }
  auto resp_or = server->ProcessRequest(req);
  CHECK(resp_or.ok(), resp_or.status().message().c_str());
  CHECK(resp_or->reply.size() == 2, "one reply per query");

  // the oracle on the same inputs
  std::vector<uint64_t> coeffs, db_ntt(params->num_pt * octx.pt_limbs());
  pir::StringEncoder enc(ep);
  if (bits_per_coeff) enc.set_bits_per_coeff(bits_per_coeff);
  for (size_t i = 0; i < params->num_pt; ++i) {
    CHECK(enc.encode(db[i], coeffs).ok(), "encode");
    orc::plain_to_ntt(octx, coeffs.data(), coeffs.size(), db_ntt.data() + i * octx.pt_limbs());
  }
  orc::GaloisKeys ogk;
  ogk.elts = gk.elts;
  ogk.data.assign(gk.limbs.begin(), gk.limbs.end());
  for (size_t qi = 0; qi < 2; ++qi) {
    std::vector<uint64_t> want;
    size_t polys = 0;
    CHECK(orc::process_query_ct(octx, rns, db_ntt.data(), params->num_pt, params->dimensions.data(), d, ogk,
                                relin.limbs.data(), req.query[qi][0].data(), 1, want, &polys) == 0, "oracle");
    const auto& reply = resp_or->reply[qi];
    CHECK(reply.size() == 1 && polys == 2 && reply[0].limbs == want, "CT-mode reply differs from the oracle");
    std::vector<uint64_t> dec(N);
    crypto.decrypt(sk, reply[0].data(), dec.data());
    auto got = enc.decode(dec, params->bytes_per_item, (*db_or)->calculate_item_offset((uint32_t)desired_index));
    CHECK(got.ok() && *got == db[desired_index], "retrieved element differs (CT mode)");
  }
  // the serialized form: relin_keys travel as a KSwitchKeys object with one slot (serialization.cpp:66-72)
  {
    pir::wire::RequestMsg m;
    for (auto& q : req.query) {
      m.query.emplace_back();
      m.query.back().ct.push_back(pir::SerializeCiphertext(ep, q[0]));
    }
    m.galois_keys = pir::SerializeGaloisKeys(ep, gk);
    pir::GaloisKeys as_slot0;
    as_slot0.elts = {pir::wire::galois_elt_of_index(0)};
    as_slot0.limbs = relin.limbs;
    const auto sp = pir::ToSealParams(ep);
    pir::wire::KSwitchKeysData R;
    R.parms_id = pir::wire::key_parms_id(sp);
    R.keys.resize(1);
    R.keys[0].resize(octx.k);
    const size_t per_digit = 2 * (octx.k + 1) * N;
    for (size_t j = 0; j < octx.k; ++j) {
      auto& c = R.keys[0][j];
      c.parms_id = R.parms_id;
      c.is_ntt_form = true;
      c.size = 2;
      c.poly_modulus_degree = N;
      c.coeff_modulus_size = octx.k + 1;
      c.limbs.assign(relin.limbs.begin() + j * per_digit, relin.limbs.begin() + (j + 1) * per_digit);
    }
    m.relin_keys = pir::wire::SaveKSwitchKeys(R);
    auto wire_resp = server->ProcessRequest(pir::wire::Serialize(m));
    CHECK(wire_resp.ok(), wire_resp.status().message().c_str());
    pir::wire::ResponseMsg back;
    CHECK(pir::wire::Parse(*wire_resp, &back) && back.reply.size() == 2 && back.reply[0].ct.size() == 1, "wire response");
    for (size_t qi = 0; qi < 2; ++qi) {
      auto ct = pir::DeserializeCiphertext(ep, back.reply[qi].ct[0]);
      CHECK(ct.ok() && ct->limbs == resp_or->reply[qi][0].limbs, "wire reply differs from the limb-level reply");
    }
  }
  if (d >= 2) {  // server.cpp:185-190 without keys: the reply keeps d + 1 polynomials
    pir::Request no_relin = req;
    no_relin.relin_keys.reset();
    no_relin.query.resize(1);
    auto r3 = server->ProcessRequest(no_relin);
    CHECK(r3.ok(), r3.status().message().c_str());
    std::vector<uint64_t> want;
    size_t polys = 0;
    CHECK(orc::process_query_ct(octx, rns, db_ntt.data(), params->num_pt, params->dimensions.data(), d, ogk, nullptr,
                                req.query[0][0].data(), 1, want, &polys) == 0, "oracle (no relin)");
    CHECK(polys == d + 1 && r3->reply[0][0].limbs == want, "size-(d+1) reply differs from the oracle");
    std::vector<uint64_t> dec(N);
    orc::decrypt_any(crypto, sk, want.data(), polys, dec.data());
    auto got = enc.decode(dec, params->bytes_per_item, (*db_or)->calculate_item_offset((uint32_t)desired_index));
    CHECK(got.ok() && *got == db[desired_index], "retrieved element differs (CT mode, no relinearization)");
  }
  std::printf("ok: CT-multiplication mode, %zu items, d=%zu, index %zu: replies bit-exact vs oracle, element recovered\n", dbsize, d,
              desired_index);
  return 0;
}

int main() {
  // error behaviour of the factories (server.cpp:37-39)
  {
    auto params = *pir::CreatePIRParameters(10, 0, 1);
    auto db_or = pir::PIRDatabase::Create(params);
    CHECK(db_or.ok(), db_or.status().message().c_str());
    auto s = pir::PIRServer::Create(*db_or, params);
    CHECK(!s.ok() && s.status().code() == PIRB_INVALID_ARGUMENT, "size mismatch must be InvalidArgument");
  }
  if (run_case(10, 0, 1, 7)) return 1;         // server_test.cpp TestProcessRequest shape
  if (run_case(82, 0, 2, 42)) return 1;        // server_test.cpp TestProcessRequest_2Dim shape (dims [10,9])
  if (run_case(1200, 64, 1, 777)) return 1;    // correctness_test.cpp:110 shape
  if (run_case(82, 0, 2, 42, 2)) return 1;     // the 2-dimensional shape, rows sharded over two contexts
  if (run_case(300, 0, 2, 123, 3)) return 1;   // dims [18,17] over three contexts
  // the reference's own CT-multiplication shapes: larger ones run out of noise budget at these parameters
  if (run_ct_case(10, 1, 7, 24, 0)) return 1;   // correctness_test.cpp:94 (d = 1)
  if (run_ct_case(9, 2, 5, 16, 10)) return 1;   // correctness_test.cpp:95 (d = 2, 16-bit t, 10 bits per coefficient)
  std::printf("SHIM_TEST_OK\n");
  return 0;
}
