// shim_bench — the reference-facing C++ call, timed like pir/cpp/benchmark.cpp:71-79 (ServerProcessRequest): one
// Request of `nq` queries per iteration through pir::PIRServer::ProcessRequest of the shim
//   e2e_cpp   ProcessRequest(const Request&)      raw-limb ciphertexts and keys in, raw-limb replies out
//   e2e_wire  ProcessRequest(const std::string&)  serialized pir.Request in, serialized pir.Response out (protobuf
//             framing + SEAL objects; keys of a returning client come from the server's cache)
// Synthetic data (uniform limbs; every kernel is data-independent).  usage:
//   shim_bench [items=4194304] [bytes=256] [dims=2] [N=4096] [plain_bits=20] [nq=8] [steps=20] [n_gpus=1]
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../pir_b200/cpp/pir_b200.hpp"

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static uint64_t splitmix(uint64_t& x) {
  uint64_t z = (x += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

int main(int argc, char** argv) {
  auto arg = [&](int i, long dflt) { return argc > i ? atol(argv[i]) : dflt; };
  const size_t items = arg(1, 1 << 22), bytes = arg(2, 256), dims = arg(3, 2), N = arg(4, 4096), bits = arg(5, 20);
  const size_t nq = arg(6, 8), steps = arg(7, 20), n_gpus = arg(8, 1);
  auto ep = pir::GenerateEncryptionParams((uint32_t)N, (uint32_t)bits);
  auto params_or = pir::CreatePIRParameters(items, bytes, dims, ep);
  if (!params_or.ok()) { std::fprintf(stderr, "params: %s\n", params_or.status().message().c_str()); return 1; }
  auto params = *params_or;
  std::vector<int> devices(n_gpus);
  for (size_t i = 0; i < n_gpus; ++i) devices[i] = (int)i;
  auto db_or = pir::PIRDatabase::Create(params, devices);
  if (!db_or.ok()) { std::fprintf(stderr, "db: %s\n", db_or.status().message().c_str()); return 1; }
  if (!(*db_or)->fill_random(2024).ok()) { std::fprintf(stderr, "fill: %s\n", pirb_last_error()); return 1; }
  auto server_or = pir::PIRServer::Create(*db_or, params, (uint32_t)std::max<size_t>(1, (nq + n_gpus - 1) / n_gpus));
  if (!server_or.ok()) { std::fprintf(stderr, "server: %s\n", server_or.status().message().c_str()); return 1; }
  auto& server = *server_or;
  const size_t k = ep.coeff_modulus.size() - 1, L = 2 * k * N;
  uint64_t seed = 99;
  pir::Request req;
  req.galois_keys.elts = pir::generate_galois_elts(N);
  req.galois_keys.limbs.resize(req.galois_keys.elts.size() * k * 2 * (k + 1) * N);
  for (size_t i = 0; i < req.galois_keys.limbs.size(); ++i)
    req.galois_keys.limbs[i] = splitmix(seed) % ep.coeff_modulus[(i / N) % (k + 1)];
  size_t dim_sum = 0;
  for (auto v : params->dimensions) dim_sum += v;
  const size_t n_ct = dim_sum / N + 1;
  req.query.assign(nq, std::vector<pir::Ciphertext>(n_ct));
  for (auto& q : req.query)
    for (auto& ct : q) {
      ct.limbs.resize(L);
      for (size_t i = 0; i < L; ++i) ct.limbs[i] = splitmix(seed) % ep.coeff_modulus[(i / N) % k];
    }
  auto time_it = [&](auto&& call, double* p50_ms) {
    for (int i = 0; i < 3; ++i) call();
    std::vector<double> lat;
    const double t0 = now_s();
    for (size_t i = 0; i < steps; ++i) {
      const double s0 = now_s();
      call();
      lat.push_back(now_s() - s0);
    }
    const double total = now_s() - t0;
    std::sort(lat.begin(), lat.end());
    *p50_ms = 1e3 * lat[lat.size() / 2];
    return (double)nq * steps / total;
  };
  bool ok = true;
  double p50_cpp = 0, p50_wire = 0;
  const double qps_cpp = time_it([&] {
    auto r = server->ProcessRequest(req);
    ok = ok && r.ok() && r->reply.size() == nq;
  }, &p50_cpp);
  // the same request on the wire
  pir::wire::RequestMsg msg;
  msg.galois_keys = pir::SerializeGaloisKeys(ep, req.galois_keys);
  for (auto& q : req.query) {
    msg.query.emplace_back();
    for (auto& ct : q) msg.query.back().ct.push_back(pir::SerializeCiphertext(ep, ct));
  }
  const std::string wire_req = pir::wire::Serialize(msg);
  size_t resp_bytes = 0;
  const double qps_wire = time_it([&] {
    auto r = server->ProcessRequest(wire_req);
    ok = ok && r.ok();
    if (r.ok()) resp_bytes = r->size();
  }, &p50_wire);
  std::printf("{\"ok\": %s, \"n_gpus\": %zu, \"queries_per_request\": %zu, \"steps\": %zu, \"num_pt\": %llu, "
              "\"e2e_cpp\": {\"value\": %.3f, \"unit\": \"queries/s\", \"p50_request_ms\": %.4f, "
              "\"what\": \"pir::PIRServer::ProcessRequest(const Request&): raw limbs in pageable std::vectors in and out, "
              "keys found in the server's cache, persistent pinned staging, graph replay\"}, "
              "\"e2e_wire\": {\"value\": %.3f, \"unit\": \"queries/s\", \"p50_request_ms\": %.4f, \"request_bytes\": %zu, "
              "\"response_bytes\": %zu, \"what\": \"pir::PIRServer::ProcessRequest(const std::string&): serialized "
              "pir.Request in, serialized pir.Response out (what pir/cpp/benchmark.cpp:71-79 times), keys of the returning "
              "client fingerprinted and found in the cache\"}, \"key_cache_hits\": %zu}\n",
              ok ? "true" : "false", n_gpus, nq, steps, (unsigned long long)params->num_pt, qps_cpp, p50_cpp, qps_wire,
              p50_wire, wire_req.size(), resp_bytes, server->key_cache_hits());
  return ok ? 0 : 1;
}
