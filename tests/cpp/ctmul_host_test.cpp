// CPU-only test of the ciphertext-multiplication mode's DEVICE code (pir_b200/csrc/pirb_behz.cuh, behz_host.h):
// the header is compiled for the host and
//   * the auxiliary bases and every base-conversion constant built by the product's host code (hm::build_behz) are
//     compared with the oracle's restatement of SEAL's RNSTool;
//   * the per-coefficient functions (base extension, scale-and-round, tensor product, relinearization mod-down) are
//     compared with the oracle on random residues;
//   * one whole upper dimension (context.cu: ct_level) is emulated launch by launch — every kernel body run for every
//     thread index of its grid (plus out-of-range ones), the NTT launches replaced by the oracle's transforms with the
//     launcher's "modulus of polynomial p = p % cycle" rule — for several queries, ragged groups, with and without a
//     relinearization key and for three-polynomial inputs, and compared limb for limb with the oracle's
//     bfv_multiply / relinearize_inplace / modular sum.
// Nothing here replaces the GPU parity tests: it pins the arithmetic and the layouts where no GPU is available.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include <cuda_runtime.h>

static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
  return (unsigned long long)(((unsigned __int128)a * b) >> 64);
}
template <typename T>
static inline T __ldg(const T* p) { return *p; }

#include "../../oracle/pir_oracle.hpp"
#include "../../pir_b200/csrc/behz_host.h"
#include "../../pir_b200/csrc/pirb_behz.cuh"

#define CHECK(cond, ...)                                       \
  do {                                                         \
    if (!(cond)) {                                             \
      std::fprintf(stderr, "FAIL %s:%d ", __FILE__, __LINE__); \
      std::fprintf(stderr, __VA_ARGS__);                       \
      std::fprintf(stderr, "\n");                              \
      return 1;                                                \
    }                                                          \
  } while (0)



static int check_constants(const orc::Context& c, const orc::RnsTool& R, const BehzC& B, const std::vector<u64>& bsk) {
  const int k = (int)c.k, nB = (int)R.nB;
  CHECK(B.k == k && B.nB == nB && (int)bsk.size() == nB + 1, "base sizes: k %d/%d nB %d/%d", B.k, k, B.nB, nB);
  for (int i = 0; i <= nB; ++i) {
    CHECK(bsk[i] == R.bsk[i].mod.q && B.bsk[i].q == bsk[i], "Bsk prime %d", i);
    CHECK(B.bsk[i].ratio_hi == R.bsk[i].mod.ratio_hi && B.bsk[i].ratio_lo == R.bsk[i].mod.ratio_lo, "Bsk ratio %d", i);
    CHECK(B.q_mod_bsk[i] == R.q_mod_bsk[i] && B.inv_mtilde_mod_bsk[i] == R.inv_mtilde_mod_bsk[i] &&
              B.inv_q_mod_bsk[i] == R.inv_q_mod_bsk[i] && B.t_mod_bsk[i] == c.t % bsk[i], "Bsk constants %d", i);
    for (int j = 0; j < k; ++j) CHECK(B.qhat_mod_bsk[i][j] == R.qhat_mod_bsk[i][j], "qhat_mod_bsk %d %d", i, j);
  }
  CHECK(bsk[nB] == R.m_sk, "m_sk last");
  for (int j = 0; j < k; ++j) {
    CHECK(B.q[j].q == c.q(j) && B.q[j].ratio_hi == c.mod(j).ratio_hi && B.q[j].ratio_lo == c.mod(j).ratio_lo, "q %d", j);
    CHECK(B.mtilde_mod_q[j] == R.mtilde_mod_q[j] && B.inv_qhat_mod_q[j] == R.inv_qhat_mod_q[j] &&
              B.qhat_mod_mtilde[j] == R.qhat_mod_mtilde[j] && B.b_mod_q[j] == R.b_mod_q[j] && B.t_mod_q[j] == c.t % c.q(j),
          "q constants %d", j);
    CHECK(B.half_P_mod_q[j] == c.half_P_mod_q[j] && B.inv_P_mod_q[j] == c.inv_P_mod_q[j], "mod-down constants %d", j);
    for (int i = 0; i < nB; ++i) CHECK(B.bhat_mod_q[j][i] == R.bhat_mod_q[j][i], "bhat_mod_q %d %d", j, i);
  }
  for (int i = 0; i < nB; ++i)
    CHECK(B.inv_bhat_mod_b[i] == R.inv_bhat_mod_b[i] && B.bhat_mod_msk[i] == R.bhat_mod_msk[i], "B constants %d", i);
  CHECK(B.neg_inv_q_mod_mtilde == R.neg_inv_q_mod_mtilde && B.inv_b_mod_msk == R.inv_b_mod_msk, "scalar constants");
  CHECK(B.P.q == c.P() && B.half_P == c.half_P, "special prime");
  return 0;
}

// the generic NTT launchers (kernels_ntt.cu) on [n_batch] x [n_polys][N] arrays: polynomial p uses table p % cycle
template <typename TableOf>
static void ntt_all(u64* data, u64 bstride, int n_batch, u64 n_polys, int cycle, size_t N, bool inverse, TableOf table_of) {
  for (int y = 0; y < n_batch; ++y)
    for (u64 p = 0; p < n_polys; ++p) {
      const orc::NttTable& T = table_of((int)(p % cycle));
      if (inverse) T.inverse((orc::u64*)data + y * bstride + p * N);
      else T.forward((orc::u64*)data + y * bstride + p * N);
    }
}

struct Case {
  int n_queries, s1;
  uint32_t n_entries, dim;
  bool relin;
};

static int run_level(const orc::Context& c, const orc::RnsTool& R, const BehzC& B, const Case& cs, std::mt19937_64& rng) {
  const size_t N = c.N, k = c.k, nb = R.n_bsk(), ctL = 2 * k * N, ptL = k * N;
  const int Q = cs.n_queries, s1 = cs.s1, sp = s1 + 1;
  const uint32_t n_entries = cs.n_entries, dim = cs.dim, n_groups = (n_entries + dim - 1) / dim;
  auto uni = [&](u64 q) { return rng() % q; };
  // inputs: lower results A [Q][n_entries][s1][k][N], selection entries S [Q] x s_bstride [dim][2][k][N], relin key
  const u64 a_bstride = (u64)n_entries * s1 * ptL, s_bstride = (u64)(dim + 3) * ctL;  // selection vectors have other entries too
  std::vector<u64> A(Q * a_bstride), S(Q * s_bstride);
  for (int y = 0; y < Q; ++y) {
    for (u64 p = 0; p < (u64)n_entries * s1; ++p)
      for (size_t j = 0; j < k; ++j)
        for (size_t n = 0; n < N; ++n) A[y * a_bstride + (p * k + j) * N + n] = uni(c.q(j));
    for (u64 p = 0; p < (u64)(dim + 3) * 2; ++p)
      for (size_t j = 0; j < k; ++j)
        for (size_t n = 0; n < N; ++n) S[y * s_bstride + (p * k + j) * N + n] = uni(c.q(j));
  }
  std::vector<u64> key(k * 2 * (k + 1) * N);
  for (size_t J = 0; J < k; ++J)
    for (size_t cc = 0; cc < 2; ++cc)
      for (size_t I = 0; I <= k; ++I)
        for (size_t n = 0; n < N; ++n) key[((J * 2 + cc) * (k + 1) + I) * N + n] = uni(c.q(I));

  // ---- emulation of ct_level (pir_b200/csrc/context.cu), one query chunk covering all queries ----
  auto grid = [&](u64 total) { return ((total + 255) / 256) * 256 + 256; };  // full blocks and one block too many
  const u64 sq_stride = (u64)dim * ctL, sb_stride = (u64)dim * 2 * nb * N;
  std::vector<u64> sq(Q * sq_stride), sb(Q * sb_stride);
  for (int y = 0; y < Q; ++y) std::memcpy(&sq[y * sq_stride], &S[y * s_bstride], sq_stride * 8);  // out-of-place NTT
  ntt_all(sq.data(), sq_stride, Q, (u64)dim * 2 * k, (int)k, N, false, [&](int j) -> const orc::NttTable& { return c.tb[j]; });
  for (int y = 0; y < Q; ++y)
    for (u64 idx = 0; idx < grid((u64)dim * 2 * N); ++idx)
      pirb::k_behz_extend_body(B, idx, y, S.data(), s_bstride, dim * 2, sb.data(), sb_stride);
  ntt_all(sb.data(), sb_stride, Q, (u64)dim * 2 * nb, (int)nb, N, false, [&](int i) -> const orc::NttTable& { return R.bsk[i]; });

  const u64 aq_s = (u64)n_entries * s1 * k * N, ab_s = (u64)n_entries * s1 * nb * N;
  const u64 dq_s = (u64)n_entries * sp * k * N, db_s = (u64)n_entries * sp * nb * N;
  const u64 dig_s = (u64)n_entries * k * (k + 1) * N, acc_s = (u64)n_entries * 2 * (k + 1) * N, x_s = (u64)n_entries * ctL;
  std::vector<u64> aq(Q * aq_s), ab(Q * ab_s), dq(Q * dq_s, ~0ull), db(Q * db_s, ~0ull), prod(Q * dq_s, ~0ull);
  for (int y = 0; y < Q; ++y) std::memcpy(&aq[y * aq_s], &A[y * a_bstride], aq_s * 8);
  ntt_all(aq.data(), aq_s, Q, (u64)n_entries * s1 * k, (int)k, N, false, [&](int j) -> const orc::NttTable& { return c.tb[j]; });
  for (int y = 0; y < Q; ++y)
    for (u64 idx = 0; idx < grid((u64)n_entries * s1 * N); ++idx)
      pirb::k_behz_extend_body(B, idx, y, A.data(), a_bstride, n_entries * s1, ab.data(), ab_s);
  ntt_all(ab.data(), ab_s, Q, (u64)n_entries * s1 * nb, (int)nb, N, false, [&](int i) -> const orc::NttTable& { return R.bsk[i]; });
  for (int y = 0; y < Q; ++y) {
    for (u64 idx = 0; idx < grid((u64)n_entries * k * N); ++idx)
      pirb::k_behz_tensor_body(B, idx, y, 0, aq.data(), aq_s, sq.data(), sq_stride, dq.data(), dq_s, n_entries, dim, s1);
    for (u64 idx = 0; idx < grid((u64)n_entries * nb * N); ++idx)
      pirb::k_behz_tensor_body(B, idx, y, 1, ab.data(), ab_s, sb.data(), sb_stride, db.data(), db_s, n_entries, dim, s1);
  }
  ntt_all(dq.data(), dq_s, Q, (u64)n_entries * sp * k, (int)k, N, true, [&](int j) -> const orc::NttTable& { return c.tb[j]; });
  ntt_all(db.data(), db_s, Q, (u64)n_entries * sp * nb, (int)nb, N, true, [&](int i) -> const orc::NttTable& { return R.bsk[i]; });
  for (int y = 0; y < Q; ++y)
    for (u64 idx = 0; idx < grid((u64)n_entries * sp * N); ++idx)
      pirb::k_behz_floor_body(B, idx, y, dq.data(), dq_s, db.data(), db_s, prod.data(), dq_s, n_entries * sp);
  const int so = cs.relin ? 2 : sp;
  const u64 out_bstride = (u64)n_groups * so * ptL;
  std::vector<u64> out(Q * out_bstride, ~0ull);
  if (cs.relin) {
    std::vector<u64> dig(Q * dig_s, ~0ull), acc(Q * acc_s, ~0ull), X(Q * x_s, ~0ull);
    for (int y = 0; y < Q; ++y)
      for (u64 idx = 0; idx < grid((u64)n_entries * k * (k + 1) * N); ++idx)
        pirb::k_relin_digits_body(B, idx, y, prod.data(), dq_s, dig.data(), dig_s, n_entries);
    ntt_all(dig.data(), dig_s, Q, (u64)n_entries * k * (k + 1), (int)k + 1, N, false,
            [&](int I) -> const orc::NttTable& { return c.tb[I]; });
    for (int y = 0; y < Q; ++y)
      for (u64 idx = 0; idx < grid((u64)n_entries * 2 * (k + 1) * N); ++idx)
        pirb::k_relin_mac_body(B, idx, y, dig.data(), dig_s, key.data(), acc.data(), acc_s, n_entries);
    ntt_all(acc.data(), acc_s, Q, (u64)n_entries * 2 * (k + 1), (int)k + 1, N, true,
            [&](int I) -> const orc::NttTable& { return c.tb[I]; });
    for (int y = 0; y < Q; ++y) {
      for (u64 idx = 0; idx < grid((u64)n_entries * 2 * k * N); ++idx)
        pirb::k_relin_finish_body(B, idx, y, prod.data(), dq_s, acc.data(), acc_s, X.data(), x_s, n_entries);
      for (u64 idx = 0; idx < grid((u64)n_groups * 2 * k * N); ++idx)
        pirb::k_ct_reduce_body(B, idx, y, X.data(), x_s, out.data(), out_bstride, n_entries, dim, 2);
    }
  } else {
    for (int y = 0; y < Q; ++y)
      for (u64 idx = 0; idx < grid((u64)n_groups * sp * k * N); ++idx)
        pirb::k_ct_reduce_body(B, idx, y, prod.data(), dq_s, out.data(), out_bstride, n_entries, dim, (uint32_t)sp);
  }

  // ---- the oracle: per query and group, sum_i relinearize(multiply(A[g*dim+i], S[i])) ----
  for (int y = 0; y < Q; ++y)
    for (uint32_t g = 0; g < n_groups; ++g) {
      std::vector<u64> want;
      for (uint32_t i = 0; i < dim && g * dim + i < n_entries; ++i) {
        const uint32_t e = g * dim + i;
        std::vector<orc::u64> pr = orc::bfv_multiply(c, R, (const orc::u64*)&A[y * a_bstride + (u64)e * s1 * ptL], s1,
                                                     (const orc::u64*)&S[y * s_bstride + (u64)i * ctL], 2);
        if (cs.relin) {
          orc::relinearize_inplace(c, pr.data(), (const orc::u64*)key.data());
          pr.resize(ctL);
        }
        if (want.empty()) {
          want.assign(pr.begin(), pr.end());
        } else {
          for (size_t x = 0; x < pr.size(); ++x) want[x] = orc::addmod(want[x], pr[x], c.mod((x / N) % k));
        }
      }
      CHECK(want.size() == (size_t)so * ptL, "oracle result size");
      const u64* got = &out[y * out_bstride + (u64)g * so * ptL];
      for (size_t x = 0; x < want.size(); ++x)
        CHECK(got[x] == want[x], "query %d group %u limb %zu: got %llu want %llu (s1=%d relin=%d)", y, g, x, got[x], want[x],
              s1, (int)cs.relin);
    }
  return 0;
}

// custom_bits != 0: three primes of that many bits (two data moduli + the special prime) instead of BFVDefault
static int run_params(uint32_t N, int plain_bits, std::mt19937_64& rng, bool full, int custom_bits = 0) {
  std::vector<orc::u64> mods;
  if (custom_bits) {
    for (orc::u64 v = ((orc::u64)1 << custom_bits) - 2 * N + 1; mods.size() < 3; v -= 2 * N)
      if (orc::is_prime_u64(v)) mods.push_back(v);
  } else {
    mods = orc::bfv_default_coeff_modulus(N);
  }
  const u64 t = orc::plain_modulus_batching(N, plain_bits);
  orc::Context c(N, mods, t);
  orc::RnsTool R(c);
  BehzC B;
  std::vector<u64> bsk;
  std::vector<u64> q(mods.begin(), mods.end() - 1);
  CHECK(pirb::hm::build_behz(q.data(), (int)c.k, mods.back(), N, c.logn, t, &B, &bsk), "build_behz failed");
  if (check_constants(c, R, B, bsk)) return 1;

  // per-coefficient functions on random and extreme residues
  const size_t k = c.k, nb = R.n_bsk();
  std::vector<orc::u64> x(k * N), yo(nb * N), dq(k * N), db(nb * N), oo(k * N);
  for (size_t j = 0; j < k; ++j)
    for (size_t n = 0; n < N; ++n) {
      x[j * N + n] = n < 4 ? (n & 1 ? c.q(j) - 1 - (n >> 1) : (n >> 1)) : rng() % c.q(j);
      dq[j * N + n] = n < 4 ? (n & 1 ? c.q(j) - 1 : 0) : rng() % c.q(j);
    }
  for (size_t i = 0; i < nb; ++i)
    for (size_t n = 0; n < N; ++n) db[i * N + n] = n < 4 ? (n & 2 ? R.bsk[i].mod.q - 1 : 0) : rng() % R.bsk[i].mod.q;
  orc::behz_extend(c, R, x.data(), yo.data());
  orc::behz_scale_and_round(c, R, dq.data(), db.data(), oo.data());
  for (size_t n = 0; n < N; ++n) {
    u64 xi[PIRB_MAX_DATA], yi[PIRB_MAX_BSK], di[PIRB_MAX_DATA], bi[PIRB_MAX_BSK], oi[PIRB_MAX_DATA];
    for (size_t j = 0; j < k; ++j) { xi[j] = x[j * N + n]; di[j] = dq[j * N + n]; }
    for (size_t i = 0; i < nb; ++i) bi[i] = db[i * N + n];
    pirb::behz_extend_coeff(B, xi, yi);
    for (size_t i = 0; i < nb; ++i) CHECK(yi[i] == yo[i * N + n], "behz_extend_coeff n=%zu i=%zu", n, i);
    pirb::behz_floor_coeff(B, di, bi, oi);
    for (size_t j = 0; j < k; ++j) CHECK(oi[j] == oo[j * N + n], "behz_floor_coeff n=%zu j=%zu", n, j);
  }
  // one upper dimension, launch by launch
  std::vector<Case> cases = {{2, 2, 5, 3, true}, {1, 2, 4, 4, false}};
  if (full) cases.push_back({1, 3, 3, 2, false});
  for (const Case& cs : cases)
    if (run_level(c, R, B, cs, rng)) return 1;
  std::printf("N=%u t=%llu: |q|=%zu |B|=%zu m_sk=%#llx — constants, per-coefficient functions and %zu emulated levels agree\n", N,
              (unsigned long long)t, k, R.nB, (unsigned long long)R.m_sk, cases.size());
  return 0;
}

int main() {
  std::mt19937_64 rng(2024);
  if (run_params(4096, 16, rng, true)) return 1;
  if (run_params(8192, 20, rng, false)) return 1;
  if (run_params(4096, 20, rng, false, 60)) return 1;  // 60-bit coefficient moduli next to the 61-bit auxiliary primes
  std::printf("CTMUL_HOST_TEST_OK\n");
  return 0;
}
