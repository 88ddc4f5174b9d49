// kernels_dist.cu — the cross-GPU exchange of the row-sharded answer path, done by the kernels themselves over
// NVLink peer memory.  No collective library is on the data path.
//
//   k_ntt_fwd_push   transform_to_ntt_inplace of the expanded selection vector (database.cpp:190,222) whose OUTPUT
//                    stores go straight into the peers' mapped exchange buffers: first-dimension entries only to the
//                    rank that owns the row, entries of the other dimensions to every rank (they multiply every row).
//   k_signal         after a producer kernel: publish a sequence number in every peer's flag block (release, system
//                    scope).  Stream order puts it behind the producer's stores.
//   k_wait           before a consumer kernel: spin (acquire, system scope) until every rank's flag has reached the
//                    sequence number; bounded by a timeout that raises an error word instead of hanging the GPU.
#include <algorithm>
#include <cstdlib>
#include <set>
#include <type_traits>
#include <utility>

#include "kernels.cuh"
#include "pirb_device.cuh"

namespace pirb {

template <int LOGN>
struct DCfg {
  static constexpr int N = 1 << LOGN;
  static constexpr int NT = (N / 8 < 512) ? N / 8 : 512;
  static constexpr size_t SMEM = sizeof(u64) * N;
};

// Persistent and deliberately NARROW: a polynomial's stores to up to n_ranks peers drain at NVLink speed (~770 GB/s
// for the whole GPU, i.e. tens of microseconds per CTA), so a full-width grid would park two 512-thread CTAs on every SM
// for most of the exchange and lock the concurrently running expansion out of registers.  A few dozen CTAs saturate
// the links; each loops over work items (polynomial p of the selection vector: entry e = p / 2k, query qi).
template <int LOGN, int ENG>
__global__ void __launch_bounds__(DCfg<LOGN>::NT, LOGN <= 13 ? 2 : 1)
k_ntt_fwd_push(const __grid_constant__ DevParams P, const u64* __restrict__ in, u64 in_qstride, u32 n_polys,
               u32 n_queries, const PushArgs A) {
  constexpr int N = DCfg<LOGN>::N, NT = DCfg<LOGN>::NT;
  extern __shared__ u64 s[];
  const int tid = threadIdx.x;
  const u32 two_k = 2 * P.k;
  for (u32 item = blockIdx.x; item < n_polys * n_queries; item += gridDim.x) {
    const u32 p = item % n_polys, qi = item / n_polys;
    const u32 e = p / two_k, within = p % two_k;
    const ModC& m = P.m[within % P.k];
    const u64* src = in + (u64)qi * in_qstride + (u64)p * N;
#pragma unroll
    for (int i = tid; i < N; i += NT) s[swz(i)] = eng_load<ENG>(src[i]);
    __syncthreads();
    eng_forward<LOGN, NT, ENG>(s, m, tid);
    // destination(s): compact per-query layout [own rows of dimension 0 | dimensions 1..] in every rank's buffer
    u32 r_lo, r_hi;
    u64 ce;  // compact entry index at the destination
    if (e < A.d0) {
      r_lo = e / A.rows_per_rank;
      r_hi = r_lo + 1;
      ce = e - r_lo * A.rows_per_rank;
    } else {
      r_lo = 0;
      r_hi = A.n_ranks;
      ce = A.rows_per_rank + (e - A.d0);
    }
    const u64 qoff = A.slot_off + (A.g_first + qi) * A.dst_qstride;
    {
      const u64 off = qoff + ce * two_k * N + (u64)within * N;
      for (int i2 = tid * 2; i2 < N; i2 += NT * 2) {
        ulonglong2 v;
        v.x = eng_store_fwd<ENG>(s[swz(i2)], m);
        v.y = eng_store_fwd<ENG>(s[swz(i2 + 1)], m);
        for (u32 r = r_lo; r < r_hi; ++r) *reinterpret_cast<ulonglong2*>(A.peers[r] + off + i2) = v;
      }
    }
    __syncthreads();  // the shared-memory polynomial is overwritten by the next item
  }
}

template <int V>
using IntC = std::integral_constant<int, V>;

cudaError_t launch_ntt_fwd_push(const DevParams& P, const u64* in, u64 in_qstride, u32 n_entries, u32 n_queries,
                                const PushArgs& A, cudaStream_t st) {
  if (!n_entries || !n_queries) return cudaSuccess;
  auto go = [&](auto ln, auto eng) -> cudaError_t {
    constexpr int LN = decltype(ln)::value;
    constexpr int EN = decltype(eng)::value;
    auto kern = k_ntt_fwd_push<LN, EN>;
    if (DCfg<LN>::SMEM > 48 * 1024) {
      static std::set<std::pair<int, const void*>> configured;
      int dev = 0;
      cudaGetDevice(&dev);
      const auto key = std::make_pair(dev, reinterpret_cast<const void*>(kern));
      if (!configured.count(key)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DCfg<LN>::SMEM);
        if (e != cudaSuccess) return e;
        configured.insert(key);
      }
    }
    const u32 n_polys = n_entries * 2 * P.k;
    static const int width = getenv("PIRB_PUSH_CTAS") ? atoi(getenv("PIRB_PUSH_CTAS")) : 64;
    const u32 grid = std::min<u64>((u64)n_polys * n_queries, (u64)std::max(1, width));
    kern<<<grid, DCfg<LN>::NT, DCfg<LN>::SMEM, st>>>(P, in, in_qstride, n_polys, n_queries, A);
    return cudaGetLastError();
  };
#define PIRB_CASE(LN)                                              \
  case LN:                                                         \
    switch (P.ntt_engine) {                                        \
      case ENG_FP64: return go(IntC<LN>{}, IntC<ENG_FP64>{});      \
      case ENG_INT_LAZY: return go(IntC<LN>{}, IntC<ENG_INT_LAZY>{}); \
      default: return go(IntC<LN>{}, IntC<ENG_INT>{});             \
    }
  switch (P.logn) {
    PIRB_CASE(11)
    PIRB_CASE(12)
    PIRB_CASE(13)
    PIRB_CASE(14)
    default: return cudaErrorInvalidValue;
  }
#undef PIRB_CASE
}

// ---------------------------------------------------------------------------------------------
// copy of the staging buffer's segments into every peer: one warp per segment chunk of 256 vectors, a lane keeps 8
// independent 16-byte loads in flight and stores each vector to all ranks (one load, n_ranks stores; consecutive lanes
// -> consecutive bytes at every peer).  No per-vector index arithmetic.
__global__ void __launch_bounds__(256)
k_push_rows(u64* const* __restrict__ peers, u32 n_ranks, const u8* __restrict__ stage, u64 src_stride, u32 seg_bytes,
            u32 n_segs, u64 dst_off, u64 dst_stride) {
  const u32 lane = threadIdx.x & 31;
  const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  const u32 vecs = seg_bytes / 16;
  const u32 chunks = (vecs + 255) / 256;
  for (u32 item = warp; item < n_segs * chunks; item += n_warps) {
    const u32 c = item / chunks, v0 = (item % chunks) * 256;
    const uint4* src = reinterpret_cast<const uint4*>(stage + (u64)c * src_stride) + v0;
    uint4 v[8];
#pragma unroll
    for (int x = 0; x < 8; ++x)
      if (v0 + lane + 32 * x < vecs) v[x] = __ldg(src + lane + 32 * x);
    for (u32 r = 0; r < n_ranks; ++r) {
      uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<u8*>(peers[r]) + dst_off + (u64)c * dst_stride) + v0;
#pragma unroll
      for (int x = 0; x < 8; ++x)
        if (v0 + lane + 32 * x < vecs) dst[lane + 32 * x] = v[x];
    }
  }
}
// first / middle dimensions of already transformed selection vectors (u64 limbs, in_qstride limbs between queries):
// entry e < d0 goes to the rank that owns row e, entries of the middle dimensions to every rank, into the compact
// per-query layout of the slot.  One block per (query, entry, chunk of 2048 vectors); same load batching.
__global__ void __launch_bounds__(256)
k_push_head(const u64* __restrict__ in, u64 in_qstride, u32 n_entries, u32 n_queries, u32 ct_vecs, const PushArgs A) {
  const u32 chunks = (ct_vecs + 2047) / 2048;
  const u32 n_items = n_queries * n_entries * chunks;
  for (u32 item = blockIdx.x; item < n_items; item += gridDim.x) {
    const u32 ch = item % chunks, qe = item / chunks;
    const u32 e = qe % n_entries, qi = qe / n_entries;
    const u32 v0 = ch * 2048 + threadIdx.x;
    const uint4* src = reinterpret_cast<const uint4*>(in + (u64)qi * in_qstride) + (u64)e * ct_vecs;
    uint4 v[8];
#pragma unroll
    for (int x = 0; x < 8; ++x)
      if (v0 + 256 * x < ct_vecs) v[x] = __ldg(src + v0 + 256 * x);
    u32 r_lo, r_hi;
    u64 ce;
    if (e < A.d0) {
      r_lo = e / A.rows_per_rank;
      r_hi = r_lo + 1;
      ce = e - r_lo * A.rows_per_rank;
    } else {
      r_lo = 0;
      r_hi = A.n_ranks;
      ce = A.rows_per_rank + (e - A.d0);
    }
    const u64 off = A.slot_off + (A.g_first + qi) * A.dst_qstride;  // limbs
    for (u32 r = r_lo; r < r_hi; ++r) {
      uint4* dst = reinterpret_cast<uint4*>(A.peers[r] + off) + ce * ct_vecs;
#pragma unroll
      for (int x = 0; x < 8; ++x)
        if (v0 + 256 * x < ct_vecs) dst[v0 + 256 * x] = v[x];
    }
  }
}
cudaError_t launch_push_head(const DevParams& P, const u64* in, u64 in_qstride, u32 n_entries, u32 n_queries,
                             const PushArgs& A, cudaStream_t st) {
  if (!n_entries || !n_queries) return cudaSuccess;
  static const int width = getenv("PIRB_PUSH_CTAS") ? atoi(getenv("PIRB_PUSH_CTAS")) : 64;
  k_push_head<<<std::max(1, width), 256, 0, st>>>(in, in_qstride, n_entries, n_queries, (u32)P.k * P.N, A);
  return cudaGetLastError();
}

cudaError_t launch_push_rows(u64* const* peers_dev, u32 n_ranks, const u8* stage, u64 src_stride, u32 seg_bytes, u32 n_segs,
                             u64 dst_off, u64 dst_stride, cudaStream_t st) {
  if (!n_segs || !seg_bytes) return cudaSuccess;
  if (seg_bytes % 16 || src_stride % 16 || dst_off % 16 || dst_stride % 16) return cudaErrorInvalidValue;
  static const int width = getenv("PIRB_PUSH_CTAS") ? atoi(getenv("PIRB_PUSH_CTAS")) : 64;
  k_push_rows<<<std::max(1, width), 256, 0, st>>>(peers_dev, n_ranks, stage, src_stride, seg_bytes, n_segs, dst_off, dst_stride);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// flags: u64 words in every rank's exchange block, flag[kind][sub][source rank]; values only grow (the step number),
// so nothing is ever reset and a late reader can never see a stale "ready".
__global__ void k_signal(u64* const* __restrict__ peers, u32 n_ranks, u64 flag_off, u64 value) {
  const u32 r = threadIdx.x;
  if (r >= n_ranks) return;
  __threadfence_system();  // everything this GPU wrote before (previous kernels of the stream included) is visible first
  u64* f = peers[r] + flag_off;
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(value) : "memory");
}
cudaError_t launch_signal(u64* const* peers_dev, u32 n_ranks, u64 flag_off_limbs, u64 value, cudaStream_t st) {
  k_signal<<<1, 32, 0, st>>>(peers_dev, n_ranks, flag_off_limbs, value);
  return cudaGetLastError();
}

__global__ void k_wait(const u64* __restrict__ flags, u32 n_ranks, u64 value, u64 timeout_ns, u64* __restrict__ err) {
  const u32 r = threadIdx.x;
  bool ok = true;
  if (r < n_ranks) {
    u64 t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      u64 v;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + r) : "memory");
      if (v >= value) break;
      u64 t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t - t0 > timeout_ns) { ok = false; break; }
      __nanosleep(200);
    }
  }
  if (!ok) atomicExch((unsigned long long*)err, value ? value : 1ull);  // the host reports it; results are void
  __threadfence_system();
}
cudaError_t launch_wait(const u64* flags, u32 n_ranks, u64 value, u64 timeout_ns, u64* err, cudaStream_t st) {
  k_wait<<<1, 32, 0, st>>>(flags, n_ranks, value, timeout_ns, err);
  return cudaGetLastError();
}

}  // namespace pirb
