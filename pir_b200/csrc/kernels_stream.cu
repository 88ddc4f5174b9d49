// kernels_stream.cu — the streaming (bandwidth-bound) kernels of the answer path:
//   k_scan            last-dimension selection-vector x database inner product (database.cpp:185-194, 238-247)
//   k_dim_mac         upper-dimension multiply_plain + add_inplace (database.cpp:229, 243-245)
//   k_ks_combine      mod-down of the key switch + SealPIR expansion butterfly (server.cpp:123-141)
//   k_mul_inv_pow_x   negacyclic shift (server.cpp:78-103)
//   k_modadd_reduce   mod-q sum of per-GPU partial replies
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "kernels.cuh"
#include "pirb_device.cuh"

namespace pirb {

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

__device__ __forceinline__ ulonglong2 ldg128(const u64* p) { return __ldg(reinterpret_cast<const ulonglong2*>(p)); }
// streaming 128-bit load that does not allocate in L1 (database limbs are read exactly once)
__device__ __forceinline__ ulonglong2 ldg128_stream(const u64* p) {
  ulonglong2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
  return r;
}

// ---------------------------------------------------------------------------------------------
// scan: grid (slice, row tile, qi*n_split + split); 128 threads x 2 limbs = 256 limbs per slice.
// Each CTA keeps 2 (polys) x R (rows) x 2 (limbs) 128-bit accumulators per thread in registers, streams
// R database tiles and the two selection-vector tiles per i1, and reduces once at the end.
// ---------------------------------------------------------------------------------------------
constexpr int SCAN_NT = 128;
constexpr int SCAN_LIMBS = 2 * SCAN_NT;

template <int R, int U, int MODE>
__global__ void __launch_bounds__(SCAN_NT, 1)  // (the explicit 1 keeps ptxas at the 124-register schedule measured at 0.766)
k_scan(const __grid_constant__ DevParams P, const u64* __restrict__ db, u64 num_pt, u32 dimL, u32 n_rows,
       const u64* __restrict__ sv, u64 sv_qstride, int n_split, u64* __restrict__ part) {
  const u32 kN = (u32)P.k * P.N;
  const u32 limb = blockIdx.x * SCAN_LIMBS + threadIdx.x * 2;
  const u32 row0 = blockIdx.y * R;
  const u32 split = blockIdx.z % n_split, qi = blockIdx.z / n_split;
  const ModC& m = P.m[limb / P.N];
  const u32 per = (dimL + n_split - 1) / n_split;
  const u32 i_lo = split * per;
  const u32 i_hi = min(dimL, i_lo + per);

  Acc<MODE> acc[R][2][2];
  const int hb = P.half_bits;
  const u64 ctL = 2ull * kN;
  const u64* svq = sv + qi * sv_qstride + limb;
  // per-row database pointers and the number of valid plaintexts in each row (short last row, database.cpp:183)
  const u64* dbr[R];
  u32 cnt[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const u64 first = (u64)(row0 + r) * dimL;
    dbr[r] = db + first * kN + limb;
    cnt[r] = (row0 + r < n_rows && first < num_pt) ? (u32)min((u64)dimL, num_pt - first) : 0;
  }
#pragma unroll 1
  for (u32 i1 = i_lo; i1 < i_hi; i1 += U) {
    ulonglong2 s0[U], s1[U], d[U][R];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const u32 i = i1 + u;
      const bool in = i < i_hi;
      s0[u] = in ? ldg128(svq + i * ctL) : make_ulonglong2(0, 0);
      s1[u] = in ? ldg128(svq + i * ctL + kN) : make_ulonglong2(0, 0);
#pragma unroll
      for (int r = 0; r < R; ++r)
        d[u][r] = (in && i < cnt[r]) ? ldg128_stream(dbr[r] + (u64)i * kN) : make_ulonglong2(0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const Opnd<MODE> a0x(s0[u].x, hb), a0y(s0[u].y, hb), a1x(s1[u].x, hb), a1y(s1[u].y, hb);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const Opnd<MODE> bx(d[u][r].x, hb), by(d[u][r].y, hb);
        acc[r][0][0].mac(a0x, bx);
        acc[r][0][1].mac(a0y, by);
        acc[r][1][0].mac(a1x, bx);
        acc[r][1][1].mac(a1y, by);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (row0 + r >= n_rows) break;
    u64* o = part + (((u64)qi * n_split + split) * n_rows + row0 + r) * ctL + limb;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      ulonglong2 v;
      v.x = acc[r][c][0].reduce(m, hb);
      v.y = acc[r][c][1].reduce(m, hb);
      *reinterpret_cast<ulonglong2*>(o + (u64)c * kN) = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// scan for a BATCH of queries sharing one pass over the database (BASELINE config 4: 64 concurrent queries).
// A CTA owns R rows x QB queries for a slice of 128 limbs (one limb per thread): every database limb it loads is
// used for QB queries x 2 ciphertext polynomials, every selection-vector limb for R rows.  Query tiles are the
// fastest grid dimension, so the CTAs that read the same database tile run together and share it through L2.
// Slices are the SLOWEST grid dimension: while one 128-limb slice is being processed its selection-vector data for
// all queries (Q x dimL x 2 KiB) stays in L2 and the slice of the database streams through exactly once.
//   grid (query tile, row tile * n_split, slice)
// ---------------------------------------------------------------------------------------------
constexpr int BATCH_NT = 128;

template <int R, int QB, int MODE>
__global__ void __launch_bounds__(BATCH_NT)
k_scan_batch(const __grid_constant__ DevParams P, const u64* __restrict__ db, u64 num_pt, u32 dimL, u32 n_rows,
             const u64* __restrict__ sv, u64 sv_qstride, int n_queries, int n_split, u64* __restrict__ part) {
  const u32 kN = (u32)P.k * P.N;
  const u64 ctL = 2ull * kN;
  const u32 limb = blockIdx.z * BATCH_NT + threadIdx.x;
  const u32 q0 = blockIdx.x * QB;
  const u32 n_row_tiles = (n_rows + R - 1) / R;
  const u32 row0 = (blockIdx.y % n_row_tiles) * R;
  const u32 split = blockIdx.y / n_row_tiles;
  const ModC& m = P.m[limb / P.N];
  const u32 per = (dimL + n_split - 1) / n_split;
  const u32 i_lo = split * per;
  const u32 i_hi = min(dimL, i_lo + per);
  const int hb = P.half_bits;
  Acc<MODE> acc[R][QB][2];
  const u64* dbr[R];
  u32 cnt[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const u64 first = (u64)(row0 + r) * dimL;
    dbr[r] = db + first * kN + limb;
    cnt[r] = (row0 + r < n_rows && first < num_pt) ? (u32)min((u64)dimL, num_pt - first) : 0;
  }
  const u64* svq[QB];
#pragma unroll
  for (int q = 0; q < QB; ++q) svq[q] = sv + (u64)min(q0 + q, (u32)n_queries - 1) * sv_qstride + limb;
#pragma unroll 1
  for (u32 i = i_lo; i < i_hi; ++i) {
    u64 s[QB][2], d[R];
#pragma unroll
    for (int q = 0; q < QB; ++q) {
      s[q][0] = __ldg(svq[q] + i * ctL);
      s[q][1] = __ldg(svq[q] + i * ctL + kN);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) d[r] = i < cnt[r] ? __ldg(dbr[r] + (u64)i * kN) : 0;
    Opnd<MODE> b0(d[0], hb);
    Opnd<MODE> b1(d[R - 1], hb);
#pragma unroll
    for (int q = 0; q < QB; ++q) {
      const Opnd<MODE> a0(s[q][0], hb), a1(s[q][1], hb);
      acc[0][q][0].mac(a0, b0);
      acc[0][q][1].mac(a1, b0);
      if (R > 1) {
        acc[R - 1][q][0].mac(a0, b1);
        acc[R - 1][q][1].mac(a1, b1);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (row0 + r >= n_rows) break;
#pragma unroll
    for (int q = 0; q < QB; ++q) {
      if (q0 + q >= (u32)n_queries) break;
      u64* o = part + (((u64)(q0 + q) * n_split + split) * n_rows + row0 + r) * ctL + limb;
      o[0] = acc[r][q][0].reduce(m, hb);
      o[kN] = acc[r][q][1].reduce(m, hb);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// batched scan, second generation (FP64 MAC only): the compute-bound regime of many queries per database pass.
// A CTA is RG row groups x 128 threads on one 128-limb slice.  Row group g keeps R rows x QB queries x 2 polynomials of
// Karatsuba accumulators in registers.  The selection tiles of the CTA's QB queries are staged ONCE per CTA in shared
// memory, already split into (low, high) halves as doubles (one 128-bit shared load per operand instead of a global
// load plus two integer->double conversions per row group), double-buffered U columns at a time with one barrier per
// chunk; each thread streams its own R database limbs through registers one chunk ahead.
//   per thread and column: 6*R*QB DFMA + 3R (database split) + 2QB (operand sums) + staging; 2QB 128-bit shared loads
//   grid (query tile, row tile * n_split, slice) — slices slowest, as for k_scan_batch
// ---------------------------------------------------------------------------------------------
template <int R, int QB, int RG, int U>
__global__ void __launch_bounds__(BATCH_NT * RG, (R * QB <= 4 && RG <= 2) ? 2 : 1)
k_scan_batch2(const __grid_constant__ DevParams P, const u64* __restrict__ db, u64 num_pt, u32 dimL, u32 n_rows,
              const u64* __restrict__ sv, u64 sv_qstride, int n_queries, int n_split, u64* __restrict__ part) {
  extern __shared__ __align__(16) double2 stage2[];  // [2][U][QB][2][128]
  constexpr int NTH = BATCH_NT * RG;
  constexpr int TILE = U * QB * 2 * BATCH_NT;  // operands per chunk
  constexpr int VPT = (TILE + NTH - 1) / NTH;
  const u32 kN = (u32)P.k * P.N;
  const u64 ctL = 2ull * kN;
  const int g = threadIdx.x / BATCH_NT, t = threadIdx.x % BATCH_NT;
  const u32 limb0 = blockIdx.z * BATCH_NT, limb = limb0 + t;
  const u32 q0 = blockIdx.x * QB;
  const u32 n_row_tiles = (n_rows + R * RG - 1) / (R * RG);
  const u32 row0 = ((blockIdx.y % n_row_tiles) * RG + g) * R;
  const u32 split = blockIdx.y / n_row_tiles;
  const ModC& m = P.m[limb / P.N];
  const u32 per = (dimL + n_split - 1) / n_split;
  const u32 i_lo = split * per;
  const u32 i_hi = min(dimL, i_lo + per);
  const int hb = P.half_bits;
  const u32 lo_mask = (1u << hb) - 1u;
  Acc<MAC_FP64> acc[R][QB][2];
  const u64* dbr[R];
  u32 cnt[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const u64 first = (u64)(row0 + r) * dimL;
    dbr[r] = db + first * kN + limb;
    cnt[r] = (row0 + r < n_rows && first < num_pt) ? (u32)min((u64)dimL, num_pt - first) : 0;
  }
  // cooperative fetch of a chunk of selection operands: operand v -> (column u, query q, polynomial p, limb l).
  // Every thread owns the same VPT operand slots for the whole loop, so their global pointers are formed once and
  // only advanced by U ciphertexts per chunk.
  const u64* sp[VPT];
  u32 su[VPT];
#pragma unroll
  for (int x = 0; x < VPT; ++x) {
    const int v = threadIdx.x + x * NTH;
    const int l = v % BATCH_NT, p = (v / BATCH_NT) % 2, q = (v / (2 * BATCH_NT)) % QB, u = v / (2 * BATCH_NT * QB);
    const u32 qq = min(q0 + q, (u32)n_queries - 1);
    su[x] = (v < TILE) ? (u32)u : 0xFFFFFFFFu;  // slots past the tile never load
    sp[x] = sv + (u64)qq * sv_qstride + (u64)(i_lo + (v < TILE ? u : 0)) * ctL + (u64)p * kN + limb0 + l;
  }
  const u64* dp[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    dp[r] = dbr[r] + (u64)i_lo * kN;
    cnt[r] = min(cnt[r], i_hi);  // one bound per row: column i is live iff i < cnt[r]
  }
  auto fetch = [&](u32 i_base, u64 (&regs)[VPT]) {
#pragma unroll
    for (int x = 0; x < VPT; ++x) {
      regs[x] = (su[x] < i_hi - min(i_base, i_hi)) ? __ldg(sp[x]) : 0ull;  // i_base + u < i_hi without overflow
      sp[x] += (u64)U * ctL;
    }
  };
  auto stash = [&](int buf, const u64 (&regs)[VPT]) {
#pragma unroll
    for (int x = 0; x < VPT; ++x) {
      const int v = threadIdx.x + x * NTH;
      if (v < TILE)
        stage2[buf * TILE + v] =
            make_double2(u32_to_double_exact((u32)regs[x] & lo_mask), u32_to_double_exact((u32)(regs[x] >> hb)));
    }
  };
  auto load_db = [&](u32 i_base, u64 (&d)[U][R]) {
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int r = 0; r < R; ++r) d[u][r] = (i_base + u < cnt[r]) ? __ldg(dp[r] + (u64)u * kN) : 0ull;
#pragma unroll
    for (int r = 0; r < R; ++r) dp[r] += (u64)U * kN;
  };
  auto compute = [&](int buf, const u64 (&d)[U][R]) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      double bl[R], bh[R], bs[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        bl[r] = u32_to_double_exact((u32)d[u][r] & lo_mask);
        bh[r] = u32_to_double_exact((u32)(d[u][r] >> hb));
        bs[r] = __dadd_rn(bl[r], bh[r]);
      }
      const double2* st = stage2 + buf * TILE + u * QB * 2 * BATCH_NT + t;
#pragma unroll
      for (int q = 0; q < QB; ++q)
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const double2 a = st[(q * 2 + p) * BATCH_NT];
          const double as = __dadd_rn(a.x, a.y);
#pragma unroll
          for (int r = 0; r < R; ++r) {
            acc[r][q][p].s0 = fma(a.x, bl[r], acc[r][q][p].s0);
            acc[r][q][p].sk = fma(as, bs[r], acc[r][q][p].sk);
            acc[r][q][p].s2 = fma(a.y, bh[r], acc[r][q][p].s2);
          }
        }
    }
  };
  // two chunks per loop trip with the roles of the register sets swapped, so nothing is copied between trips
  u64 svr[VPT], da[U][R], dbq[U][R];
  fetch(i_lo, svr);
  load_db(i_lo, da);
  stash(0, svr);
  __syncthreads();
#pragma unroll 1
  for (u32 i1 = i_lo; i1 < i_hi; i1 += 2 * U) {
    fetch(i1 + U, svr);  // the next chunk's operands travel while this chunk is multiplied
    load_db(i1 + U, dbq);
    compute(0, da);
    stash(1, svr);
    __syncthreads();
    if (i1 + U >= i_hi) break;
    fetch(i1 + 2 * U, svr);
    load_db(i1 + 2 * U, da);
    compute(1, dbq);
    stash(0, svr);
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (row0 + r >= n_rows) break;
#pragma unroll
    for (int q = 0; q < QB; ++q) {
      if (q0 + q >= (u32)n_queries) break;
      u64* o = part + (((u64)(q0 + q) * n_split + split) * n_rows + row0 + r) * ctL + limb;
      o[0] = acc[r][q][0].reduce(m, hb);
      o[kN] = acc[r][q][1].reduce(m, hb);
    }
  }
}

// tuning knobs (overridable through the environment for sweeps): rows per CTA, unroll, generic MAC
static void scan_pick(const DevParams& P, u32 n_rows, int* R, int* U, int* mode) {
  // measured on B200 (profiles/): 2 rows x unroll 2 with the FP64 MAC is the fastest shape
  int r = n_rows >= 2 ? 2 : 1;
  int u = r == 2 ? 2 : 4;
  r = env_int("PIRB_SCAN_R", r);
  u = env_int("PIRB_SCAN_U", u);
  *R = r; *U = u;
  *mode = std::min(P.mac_mode, env_int("PIRB_MAC_MODE", 2));
}

void scan_config(const DevParams& P, u32 dimL, u32 n_rows, int n_queries, int sm_count, int* n_split) {
  // The grid is (slices x row tiles x queries) CTAs, each optionally working on 1/n_split of the last dimension.  Small
  // and medium databases give only one to three waves of CTAs, so the last, partly filled wave is a large share of the
  // kernel (2.27 waves = 32 % tail on BASELINE config 2 and on a 1/8 shard of config 4).  Choose the split whose CTA
  // count fills whole waves best, charging the extra partial-result traffic (every split writes its own row results,
  // which the inverse NTT's load side sums) and keeping at least 8 database tiles per CTA.
  int R, U, mode;
  scan_pick(P, n_rows, &R, &U, &mode);
  const u32 slices = (u32)P.k * P.N / SCAN_LIMBS;
  const u64 row_tiles = (n_rows + R - 1) / R;
  const u64 base = (u64)slices * row_tiles * n_queries;
  const u64 slots = (u64)sm_count * env_int("PIRB_SCAN_CTAS_PER_SM", 4);  // resident CTAs (124 registers x 128 threads)
  // exact lazy accumulation chains have a maximum length (pirb_device.cuh)
  const u32 max_terms = mode >= MAC_FP64 ? P.mac_max_terms : (mode == MAC_INT24 ? PIRB_SMALL_MAX_TERMS : P.wide_max_terms);
  const int min_split = std::max(1, (int)((dimL + max_terms - 1) / max_terms));
  const int max_split = std::max(min_split, (int)(dimL / 8));
  const double db_bytes = (double)row_tiles * R * dimL + 2.0 * dimL;  // in plaintext-sized units (x pt bytes)
  int best = min_split;
  double best_score = -1.0;
  // six or more waves without splitting: the tail is already small (measured: 76 % of the copy bandwidth on the 6.7 GiB
  // database), keep one pass per row
  for (int sp = min_split; sp <= max_split && sp <= 64 && base < 6 * slots; ++sp) {
    const double ctas = (double)base * sp;
    const double waves = std::ceil(ctas / (double)slots);
    const double fill = ctas / ((double)slots * waves);               // 1.0 = every wave full
    const double extra = 2.0 * 2.0 * (double)n_rows * (sp - 1);       // partial rows written + read again (2 pt each)
    const double score = fill / (1.0 + extra / db_bytes) * (waves >= 2.0 || ctas <= (double)slots ? 1.0 : 0.9);
    if (score > best_score + 1e-9) { best_score = score; best = sp; }
  }
  int sp = env_int("PIRB_SCAN_SPLIT", best);
  if (sp < min_split) sp = min_split;
  *n_split = sp;
}

cudaError_t launch_scan(const DevParams& P, const u64* db, u64 num_pt, u32 dimL, u32 n_rows, const u64* sv,
                        u64 sv_qstride, int n_queries, int n_split, u64* part, cudaStream_t st) {
  if (!n_rows || !n_queries) return cudaSuccess;
  int R, U, mode;
  scan_pick(P, n_rows, &R, &U, &mode);
  {
    const u32 chain = (dimL + n_split - 1) / n_split;
    if (mode >= MAC_FP64 && chain > P.mac_max_terms) mode = P.mac_mode >= 1 && P.half_bits <= 24 ? MAC_INT24 : MAC_WIDE;
    if (mode == MAC_INT24 && chain > PIRB_SMALL_MAX_TERMS) mode = MAC_WIDE;
    if (mode == MAC_WIDE && chain > P.wide_max_terms) return cudaErrorInvalidValue;  // scan_config splits finer than this
  }
  const u32 slices = (u32)P.k * P.N / SCAN_LIMBS;
  dim3 grid(slices, (n_rows + R - 1) / R, n_queries * n_split);
  if (grid.y > 65535 || grid.z > 65535) return cudaErrorInvalidConfiguration;
  if (n_queries >= env_int("PIRB_SCAN_BATCH_MIN", 4)) {
    // batch of queries on CUDA cores (shards too small for the tensor-core scan, or moduli it does not cover): share
    // every database tile between QB queries
    if (mode == MAC_FP64) {
      const int R2 = n_rows >= 2 ? 2 : 1;
      int RG2 = 4;
      while (RG2 > 1 && (u32)(R2 * RG2) > n_rows + R2 - 1) RG2 >>= 1;  // do not idle whole row groups on small shards
      const u32 tiles = (n_rows + R2 * RG2 - 1) / (R2 * RG2);
      dim3 g2((n_queries + 1) / 2, tiles * n_split, (u32)P.k * P.N / BATCH_NT);
      if (g2.y > 65535 || g2.z > 65535) return cudaErrorInvalidConfiguration;
      const size_t smem2 = (size_t)2 * 2 * 2 * 2 * BATCH_NT * sizeof(double2);
#define B2_CASE(RR, GG)                                                                                            \
  if (R2 == RR && RG2 == GG) {                                                                                     \
    k_scan_batch2<RR, 2, GG, 2><<<g2, BATCH_NT * GG, smem2, st>>>(P, db, num_pt, dimL, n_rows, sv, sv_qstride,     \
                                                                  n_queries, n_split, part);                      \
    return cudaGetLastError();                                                                                     \
  }
      // measured best shape (2 rows x 2 queries per thread, 4 row groups, 2 columns per barrier) and its fallbacks
      B2_CASE(2, 4) B2_CASE(2, 2) B2_CASE(2, 1) B2_CASE(1, 4) B2_CASE(1, 2) B2_CASE(1, 1)
#undef B2_CASE
    }
    const int QB = n_queries >= 4 ? 4 : 2;
    const int RB = n_rows >= 2 ? 2 : 1;
    dim3 bgrid((n_queries + QB - 1) / QB, ((n_rows + RB - 1) / RB) * n_split, (u32)P.k * P.N / BATCH_NT);
    if (bgrid.y > 65535 || bgrid.z > 65535) return cudaErrorInvalidConfiguration;
#define BATCH_CASE(RR, QQ)                                                                                         \
  if (RB == RR && QB == QQ) {                                                                                      \
    if (mode == MAC_FP64) k_scan_batch<RR, QQ, MAC_FP64><<<bgrid, BATCH_NT, 0, st>>>(P, db, num_pt, dimL, n_rows, sv, sv_qstride, n_queries, n_split, part); \
    else if (mode == MAC_INT24) k_scan_batch<RR, QQ, MAC_INT24><<<bgrid, BATCH_NT, 0, st>>>(P, db, num_pt, dimL, n_rows, sv, sv_qstride, n_queries, n_split, part); \
    else k_scan_batch<RR, QQ, MAC_WIDE><<<bgrid, BATCH_NT, 0, st>>>(P, db, num_pt, dimL, n_rows, sv, sv_qstride, n_queries, n_split, part); \
    return cudaGetLastError();                                                                                     \
  }
    BATCH_CASE(1, 2) BATCH_CASE(1, 4) BATCH_CASE(2, 2) BATCH_CASE(2, 4)
#undef BATCH_CASE
  }
#define SCAN_CASE(RR, UU)                                                                                          \
  if (R == RR && U == UU) {                                                                                        \
    if (mode == MAC_FP64) k_scan<RR, UU, MAC_FP64><<<grid, SCAN_NT, 0, st>>>(P, db, num_pt, dimL, n_rows, sv, sv_qstride, n_split, part); \
    else if (mode == MAC_INT24) k_scan<RR, UU, MAC_INT24><<<grid, SCAN_NT, 0, st>>>(P, db, num_pt, dimL, n_rows, sv, sv_qstride, n_split, part); \
    else k_scan<RR, UU, MAC_WIDE><<<grid, SCAN_NT, 0, st>>>(P, db, num_pt, dimL, n_rows, sv, sv_qstride, n_split, part);      \
    return cudaGetLastError();                                                                                     \
  }
  SCAN_CASE(1, 1) SCAN_CASE(1, 2) SCAN_CASE(1, 4)
  SCAN_CASE(2, 1) SCAN_CASE(2, 2) SCAN_CASE(2, 4)
#undef SCAN_CASE
  return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------------------------
// upper-dimension MAC: grid (slice, x = output ct of the group, (qi*n_groups + g)*n_split + split)
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(SCAN_NT)
k_dim_mac(const __grid_constant__ DevParams P, const u64* __restrict__ pts, u64 pts_qstride,
          const u64* __restrict__ sv, u64 sv_qstride, u32 dim, u32 n_entries_in, u32 n_groups, u32 w_out, int n_split,
          u64* __restrict__ part) {
  const u32 kN = (u32)P.k * P.N;
  const u64 ctL = 2ull * kN;
  const u32 limb = blockIdx.x * SCAN_LIMBS + threadIdx.x * 2;
  const u32 x = blockIdx.y;
  const u32 split = blockIdx.z % n_split;
  const u32 gq = blockIdx.z / n_split;
  const u32 g = gq % n_groups, qi = gq / n_groups;
  const ModC& m = P.m[limb / P.N];
  const u32 cnt = min(dim, n_entries_in - g * dim);
  const u32 per = (dim + n_split - 1) / n_split;
  const u32 i_lo = split * per, i_hi = min(cnt, i_lo + per);
  const int hb = P.half_bits;
  Acc<MODE> acc[2][2];
  const u64* svq = sv + qi * sv_qstride + limb;
  const u64* pq = pts + qi * pts_qstride + ((u64)g * dim * w_out + x) * kN + limb;
#pragma unroll 4
  for (u32 i = i_lo; i < i_hi; ++i) {
    const ulonglong2 s0 = ldg128(svq + i * ctL);
    const ulonglong2 s1 = ldg128(svq + i * ctL + kN);
    const ulonglong2 d = ldg128_stream(pq + (u64)i * w_out * kN);
    const Opnd<MODE> bx(d.x, hb), by(d.y, hb);
    acc[0][0].mac(Opnd<MODE>(s0.x, hb), bx);
    acc[0][1].mac(Opnd<MODE>(s0.y, hb), by);
    acc[1][0].mac(Opnd<MODE>(s1.x, hb), bx);
    acc[1][1].mac(Opnd<MODE>(s1.y, hb), by);
  }
  u64* o = part + ((((u64)qi * n_split + split) * n_groups + g) * w_out + x) * ctL + limb;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    ulonglong2 v;
    v.x = acc[c][0].reduce(m, hb);
    v.y = acc[c][1].reduce(m, hb);
    *reinterpret_cast<ulonglong2*>(o + (u64)c * kN) = v;
  }
}

static u32 mac_chain_limit(const DevParams& P, int mode) {
  return mode >= MAC_FP64 ? P.mac_max_terms : (mode == MAC_INT24 ? PIRB_SMALL_MAX_TERMS : P.wide_max_terms);
}

// split of an upper dimension's sum over `len` entries: enough CTAs to fill the GPU, and never a lazy accumulation
// chain longer than the exact range of the context's multiply-accumulate mode (pirb_device.cuh)
int dim_mac_config(const DevParams& P, u64 base_ctas, u32 len, int sm_count) {
  const u64 want = (u64)sm_count * 4;
  int s = 1;
  if (base_ctas < want) s = (int)((want + base_ctas - 1) / base_ctas);
  const int max_split = (int)((len + 3) / 4);
  if (s > max_split) s = max_split;
  const int mode = std::min(P.mac_mode, env_int("PIRB_MAC_MODE", 2));
  const u32 max_terms = mac_chain_limit(P, mode);
  const int min_split = (int)((len + max_terms - 1) / max_terms);
  if (s < min_split) s = min_split;
  return s < 1 ? 1 : s;
}

cudaError_t launch_dim_mac(const DevParams& P, const u64* pts, u64 pts_qstride, const u64* sv, u64 sv_qstride,
                           int n_queries, u32 dim, u32 n_entries_in, u32 n_groups, u32 w_out, int n_split, u64* part,
                           cudaStream_t st) {
  if (!n_groups || !n_queries) return cudaSuccess;
  const u32 slices = (u32)P.k * P.N / SCAN_LIMBS;
  dim3 grid(slices, w_out, (unsigned)n_queries * n_groups * n_split);
  if (grid.y > 65535 || grid.z > 65535) return cudaErrorInvalidConfiguration;
  const int mode = std::min(P.mac_mode, env_int("PIRB_MAC_MODE", 2));
  if ((dim + n_split - 1) / n_split > mac_chain_limit(P, mode)) return cudaErrorInvalidValue;
  if (mode == MAC_FP64)
    k_dim_mac<MAC_FP64><<<grid, SCAN_NT, 0, st>>>(P, pts, pts_qstride, sv, sv_qstride, dim, n_entries_in, n_groups, w_out, n_split, part);
  else if (mode == MAC_INT24)
    k_dim_mac<MAC_INT24><<<grid, SCAN_NT, 0, st>>>(P, pts, pts_qstride, sv, sv_qstride, dim, n_entries_in, n_groups, w_out, n_split, part);
  else
    k_dim_mac<MAC_WIDE><<<grid, SCAN_NT, 0, st>>>(P, pts, pts_qstride, sv, sv_qstride, dim, n_entries_in, n_groups, w_out, n_split, part);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// key-switch mod-down + expansion combine.  grid (node z, N/256, j); one thread per coefficient.
//   md_c = (acc_c[j] - ((acc_c[P] + P/2 mod P) mod q_j - (P/2 mod q_j))) * P^{-1}  mod q_j
//   c0 = (sigma_g(p0) + md_0, md_1)
//   mode 0: dst[kk] = src + c0 ; dst[kk + 2^j] = (src - c0) * x^{-2^j}           (server.cpp:123-141)
//   mode 1: dst[kk] = c0                                                          (substitute_power_x_inplace)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 mod_down(u64 a, u64 last, const DevParams& P, int j, u64 Pq) {
  const ModC& m = P.m[j];
  u64 l = last + P.half_P;
  l = l >= Pq ? l - Pq : l;
  u64 r = submod(barrett64(l, m.q, m.ratio_hi), P.half_P_mod[j], m.q);
  return shoup(submod(a, r, m.q), P.inv_P[j], P.inv_P_s[j], m.q);
}

__global__ void __launch_bounds__(256)
k_ks_combine(const __grid_constant__ DevParams P, u64* __restrict__ work, const LevelArgs L,
             const u64* __restrict__ acc, int mode) {
  const u32 N = P.N;
  const int k = P.k;
  const u32 z = blockIdx.x;
  const u32 n = blockIdx.y * 256 + threadIdx.x;
  const int j = blockIdx.z;
  const u32 kk = z & ((1u << L.j) - 1);
  const u32 tq = z >> L.j;
  const u32 ti = tq % L.n_trees, qi = tq / L.n_trees;
  const u64 ctL = 2ull * k * N;
  const u64 q = P.m[j].q, Pq = P.m[k].q;
  const u64* src = work + qi * L.q_stride + L.src_off[ti] + kk * ctL;
  u64* dstE = work + qi * L.q_stride + L.dst_off[ti] + kk * ctL;
  const u64* a = acc + (u64)z * 2 * (k + 1) * N;
  const u64 md0 = mod_down(a[(u64)j * N + n], a[(u64)k * N + n], P, j, Pq);
  const u64 md1 = mod_down(a[(u64)(k + 1 + j) * N + n], a[(u64)(k + 1 + k) * N + n], P, j, Pq);
  const u64 c00 = addmod(galois_gather(src + (u64)j * N, n, L.ginv, N, q), md0, q);
  const u64 c01 = md1;
  if (mode == 1) {
    dstE[(u64)j * N + n] = c00;
    dstE[(u64)(k + j) * N + n] = c01;
    return;
  }
  const u64 p0 = src[(u64)j * N + n], p1 = src[(u64)(k + j) * N + n];
  dstE[(u64)j * N + n] = addmod(p0, c00, q);
  dstE[(u64)(k + j) * N + n] = addmod(p1, c01, q);
  u64* dstO = dstE + ((u64)ctL << L.j);
  const u32 r = n + ((2 * N - (1u << L.j)) & (2 * N - 1));
  const u32 idx = r & (N - 1);
  u64 d0 = submod(p0, c00, q), d1 = submod(p1, c01, q);
  if (r & N) { d0 = negmod(d0, q); d1 = negmod(d1, q); }
  dstO[(u64)j * N + idx] = d0;
  dstO[(u64)(k + j) * N + idx] = d1;
}

cudaError_t launch_ks_combine(const DevParams& P, u64* work, const LevelArgs& L, const u64* acc, int mode,
                              cudaStream_t st) {
  const unsigned nodes = (unsigned)L.n_queries * L.n_trees << L.j;
  if (!nodes) return cudaSuccess;
  k_ks_combine<<<dim3(nodes, P.N / 256, P.k), 256, 0, st>>>(P, work, L, acc, mode);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_mul_inv_pow_x(const __grid_constant__ DevParams P, const u64* __restrict__ in, u64* __restrict__ out, u32 shift) {
  const u32 N = P.N;
  const u32 n = blockIdx.y * 256 + threadIdx.x;
  const u32 poly = blockIdx.x;  // ct*2k + i*k + j
  const u64 q = P.m[poly % P.k].q;
  const u32 r = n + shift;
  u64 v = in[(u64)poly * N + n];
  if (r & N) v = negmod(v, q);
  out[(u64)poly * N + (r & (N - 1))] = v;
}
cudaError_t launch_mul_inv_pow_x(const DevParams& P, const u64* in, u64* out, u32 kpow, int n_cts, cudaStream_t st) {
  if (n_cts <= 0) return cudaSuccess;
  const u32 twoN = 2 * P.N;
  const u32 shift = (twoN - kpow) % twoN;  // server.cpp:87-88 (uint32 arithmetic, wraps like the reference)
  k_mul_inv_pow_x<<<dim3(n_cts * 2 * P.k, P.N / 256), 256, 0, st>>>(P, in, out, shift);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_modadd_reduce(const __grid_constant__ DevParams P, const u64* __restrict__ in, u64 stride, int n_parts,
                u64* __restrict__ out, u64 total_limbs, u64 in_bstride, u64 out_bstride) {
  const u64 i = ((u64)blockIdx.x * 256 + threadIdx.x) * 2;
  if (i >= total_limbs) return;
  in += blockIdx.y * in_bstride;
  out += blockIdx.y * out_bstride;
  const u64 q = P.m[(i / P.N) % P.k].q;
  ulonglong2 v = ldg128(in + i);
  for (int g = 1; g < n_parts; ++g) {
    const ulonglong2 w = ldg128(in + g * stride + i);
    v.x = addmod(v.x, w.x, q);
    v.y = addmod(v.y, w.y, q);
  }
  *reinterpret_cast<ulonglong2*>(out + i) = v;
}
cudaError_t launch_modadd_reduce(const DevParams& P, const u64* in, u64 stride, int n_parts, u64* out, u64 n_cts,
                                 cudaStream_t st, int n_batch, u64 in_bstride, u64 out_bstride) {
  const u64 total = n_cts * 2 * P.k * P.N;
  if (!total || n_batch <= 0) return cudaSuccess;
  k_modadd_reduce<<<dim3((unsigned)((total / 2 + 255) / 256), (unsigned)n_batch), 256, 0, st>>>(P, in, stride, n_parts, out,
                                                                                              total, in_bstride, out_bstride);
  return cudaGetLastError();
}

// Same reduction, but every partial is read through its own pointer: with CUDA IPC / peer access those are
// other GPUs' buffers, so the loads travel over NVLink inside the reducing kernel (no staging copy).
__global__ void __launch_bounds__(256)
k_modadd_reduce_ptrs(const __grid_constant__ DevParams P, const u64* const* __restrict__ peers, int n_parts,
                     u64 offset, u64* __restrict__ out, u64 total_limbs) {
  const u64 i = ((u64)blockIdx.x * 256 + threadIdx.x) * 2;
  if (i >= total_limbs) return;
  const u64 q = P.m[(i / P.N) % P.k].q;
  ulonglong2 v = *reinterpret_cast<const ulonglong2*>(peers[0] + offset + i);
  for (int g = 1; g < n_parts; ++g) {
    const ulonglong2 w = *reinterpret_cast<const ulonglong2*>(peers[g] + offset + i);
    v.x = addmod(v.x, w.x, q);
    v.y = addmod(v.y, w.y, q);
  }
  *reinterpret_cast<ulonglong2*>(out + i) = v;
}
cudaError_t launch_modadd_reduce_ptrs(const DevParams& P, const u64* const* peers_dev, int n_parts, u64 offset_limbs,
                                      u64* out, u64 n_cts, cudaStream_t st) {
  const u64 total = n_cts * 2 * P.k * P.N;
  if (!total) return cudaSuccess;
  k_modadd_reduce_ptrs<<<(unsigned)((total / 2 + 255) / 256), 256, 0, st>>>(P, peers_dev, n_parts, offset_limbs, out,
                                                                            total);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_place_roots(const u64* __restrict__ query, u64 query_qstride, u64* __restrict__ work,
              const u64* __restrict__ root_off, int n_trees, u64 q_stride, u64 ctL) {
  const u32 tq = blockIdx.y;
  const u32 ti = tq % n_trees, qi = tq / n_trees;
  const u64 i = ((u64)blockIdx.x * 256 + threadIdx.x) * 2;
  if (i >= ctL) return;
  const ulonglong2 v = ldg128(query + qi * query_qstride + (u64)ti * ctL + i);
  *reinterpret_cast<ulonglong2*>(work + qi * q_stride + root_off[ti] + i) = v;
}
cudaError_t launch_place_roots(const DevParams& P, const u64* query, u64 query_qstride, u64* work, const u64* root_off,
                               int n_trees, int n_queries, u64 q_stride, cudaStream_t st) {
  if (!n_trees || !n_queries) return cudaSuccess;
  const u64 ctL = 2ull * P.k * P.N;
  k_place_roots<<<dim3((unsigned)((ctL / 2 + 255) / 256), n_trees * n_queries), 256, 0, st>>>(
      query, query_qstride, work, root_off, n_trees, q_stride, ctL);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// StringEncoder::encode on the device (string_encoder.cpp:58-80, 108-122): coefficient c of plaintext p holds bits
// [c*b, (c+1)*b) (MSB first) of the concatenation of the plaintext's items, zero-padded.  One thread per coefficient.
__global__ void __launch_bounds__(256)
k_pack_items(const u8* __restrict__ bytes, u64* __restrict__ coeffs, u32 N, u32 bits, u64 bytes_per_pt,
             u64 total_bytes) {
  const u64 pt = blockIdx.y;
  const u32 c = blockIdx.x * 256 + threadIdx.x;
  if (c >= N) return;
  const u64 base = pt * bytes_per_pt;
  const u64 avail = total_bytes > base ? min(bytes_per_pt, total_bytes - base) : 0;  // the last plaintext may be short
  const u64 bit0 = (u64)c * bits;
  u64 v = 0;
  // gather up to 9 bytes covering the (at most 32) wanted bits
  const u64 byte0 = bit0 >> 3;
  const u32 skip = (u32)(bit0 & 7);
  unsigned __int128 acc = 0;
  const u32 nbytes = (skip + bits + 7) >> 3;
  for (u32 i = 0; i < nbytes; ++i) {
    const u64 idx = byte0 + i;
    const u8 b = idx < avail ? bytes[base + idx] : 0;
    acc = (acc << 8) | b;
  }
  const u32 tail = nbytes * 8 - skip - bits;
  v = (u64)((acc >> tail) & (((unsigned __int128)1 << bits) - 1));
  coeffs[pt * N + c] = v;
}
cudaError_t launch_pack_items(const u8* bytes, u64* coeffs, u32 N, u32 bits, u64 bytes_per_pt, u64 total_bytes,
                              u64 n_pt, cudaStream_t st) {
  if (!n_pt) return cudaSuccess;
  k_pack_items<<<dim3((N + 255) / 256, (unsigned)n_pt), 256, 0, st>>>(bytes, coeffs, N, bits, bytes_per_pt, total_bytes);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// L2 warm-up of constants (NTT tables, the client's Galois keys): one prefetch per 128-byte line, fire and forget.
__global__ void __launch_bounds__(256) k_prefetch_l2(const char* __restrict__ p, u64 lines) {
  const u64 l = (u64)blockIdx.x * 256 + threadIdx.x;
  if (l < lines) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + l * 128));
}
cudaError_t launch_prefetch_l2(const void* p, u64 bytes, cudaStream_t st) {
  const u64 lines = (bytes + 127) / 128;
  if (!lines) return cudaSuccess;
  k_prefetch_l2<<<(unsigned)((lines + 255) / 256), 256, 0, st>>>(reinterpret_cast<const char*>(p), lines);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_fill_random(const __grid_constant__ DevParams P, u64* __restrict__ out, int cycle, int off, u64 seed, u64 poly0) {
  const u64 poly = blockIdx.x;
  const u64 q = P.m[(poly % cycle) + off].q;
  for (u32 n = threadIdx.x; n < P.N; n += 256) {
    // counter = GLOBAL polynomial index: a row shard holds exactly the limbs the unsharded database holds there
    u64 x = seed + ((poly0 + poly) * P.N + n) * 0x9e3779b97f4a7c15ULL;  // splitmix64 finaliser as a counter-based hash
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
    x ^= x >> 31;
    out[poly * P.N + n] = x % q;
  }
}
cudaError_t launch_fill_random(const DevParams& P, u64* out, u64 n_polys, int cycle, int off, u64 seed, u64 poly0,
                               cudaStream_t st) {
  if (!n_polys) return cudaSuccess;
  k_fill_random<<<(unsigned)n_polys, 256, 0, st>>>(P, out, cycle, off, seed, poly0);
  return cudaGetLastError();
}

}  // namespace pirb
