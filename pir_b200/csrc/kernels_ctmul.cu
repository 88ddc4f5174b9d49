// kernels_ctmul.cu — the ciphertext-multiplication mode of the upper dimensions (database.cpp:202-211):
//   Evaluator::multiply(lower_result, selection_ct)  +  Evaluator::relinearize_inplace   [SEAL 3.5.6, BEHZ]
// Outside the NTTs (kernels_ntt.cu, run with a second table set for the auxiliary base Bsk) every step is independent
// per coefficient; the arithmetic lives in pirb_behz.cuh, these kernels add the indexing.  One thread per coefficient,
// consecutive threads on consecutive coefficients: every access is a coalesced stream.  Batch index (a query of the
// request) = blockIdx.y, *_bstride limbs apart.
//
// This mode is off by default in the reference and outside the bandwidth-bound path the rest of this library is built
// around: the kernels are written for exactness and plain streaming, not tuned.
#include "kernels.cuh"
#include "pirb_behz.cuh"

namespace pirb {

namespace {
constexpr int CT_NT = 256;
inline unsigned ct_blocks(u64 total) { return (unsigned)((total + CT_NT - 1) / CT_NT); }
}  // namespace

// The kernel bodies (index arithmetic + the per-coefficient functions) live in pirb_behz.cuh as k_*_body(B, idx, batch, ...)
// so that the host test can run them thread by thread; a kernel is its body at idx = global thread index, batch = blockIdx.y.
__global__ void __launch_bounds__(CT_NT)
k_behz_extend(const __grid_constant__ BehzC B, const u64* __restrict__ in, u64 in_bstride, u32 n_polys,
              u64* __restrict__ out, u64 out_bstride) {
  k_behz_extend_body(B, (u64)blockIdx.x * CT_NT + threadIdx.x, blockIdx.y, in, in_bstride, n_polys, out, out_bstride);
}
__global__ void __launch_bounds__(CT_NT)
k_behz_tensor(const __grid_constant__ BehzC B, int base, const u64* __restrict__ A, u64 a_bstride,
              const u64* __restrict__ S, u64 s_bstride, u64* __restrict__ D, u64 d_bstride, u32 n_entries, u32 dim,
              int s1) {
  k_behz_tensor_body(B, (u64)blockIdx.x * CT_NT + threadIdx.x, blockIdx.y, base, A, a_bstride, S, s_bstride, D,
                     d_bstride, n_entries, dim, s1);
}
__global__ void __launch_bounds__(CT_NT)
k_behz_floor(const __grid_constant__ BehzC B, const u64* __restrict__ Dq, u64 dq_bstride, const u64* __restrict__ Db,
             u64 db_bstride, u64* __restrict__ out, u64 out_bstride, u32 n_polys) {
  k_behz_floor_body(B, (u64)blockIdx.x * CT_NT + threadIdx.x, blockIdx.y, Dq, dq_bstride, Db, db_bstride, out,
                    out_bstride, n_polys);
}
__global__ void __launch_bounds__(CT_NT)
k_relin_digits(const __grid_constant__ BehzC B, const u64* __restrict__ prod, u64 p_bstride, u64* __restrict__ dig,
               u64 dig_bstride, u32 n_entries) {
  k_relin_digits_body(B, (u64)blockIdx.x * CT_NT + threadIdx.x, blockIdx.y, prod, p_bstride, dig, dig_bstride,
                      n_entries);
}
__global__ void __launch_bounds__(CT_NT)
k_relin_mac(const __grid_constant__ BehzC B, const u64* __restrict__ dig, u64 dig_bstride,
            const u64* __restrict__ key, u64* __restrict__ acc, u64 acc_bstride, u32 n_entries) {
  k_relin_mac_body(B, (u64)blockIdx.x * CT_NT + threadIdx.x, blockIdx.y, dig, dig_bstride, key, acc, acc_bstride,
                   n_entries);
}
__global__ void __launch_bounds__(CT_NT)
k_relin_finish(const __grid_constant__ BehzC B, const u64* __restrict__ prod, u64 p_bstride,
               const u64* __restrict__ acc, u64 acc_bstride, u64* __restrict__ X, u64 x_bstride, u32 n_entries) {
  k_relin_finish_body(B, (u64)blockIdx.x * CT_NT + threadIdx.x, blockIdx.y, prod, p_bstride, acc, acc_bstride, X,
                      x_bstride, n_entries);
}
__global__ void __launch_bounds__(CT_NT)
k_ct_reduce(const __grid_constant__ BehzC B, const u64* __restrict__ X, u64 x_bstride, u64* __restrict__ out,
            u64 out_bstride, u32 n_entries, u32 dim, u32 polys) {
  k_ct_reduce_body(B, (u64)blockIdx.x * CT_NT + threadIdx.x, blockIdx.y, X, x_bstride, out, out_bstride, n_entries,
                   dim, polys);
}

// ---------------------------------------------------------------------------------------------
#define CT_LAUNCH(total, n_batch, kern, ...)                                  \
  do {                                                                        \
    if ((total) == 0 || (n_batch) <= 0) return cudaSuccess;                   \
    kern<<<dim3(ct_blocks(total), (unsigned)(n_batch)), CT_NT, 0, st>>>(__VA_ARGS__); \
    return cudaGetLastError();                                                \
  } while (0)

cudaError_t launch_behz_extend(const BehzC& B, const u64* in, u64 in_bstride, u32 n_polys, u64* out, u64 out_bstride,
                               int n_batch, cudaStream_t st) {
  CT_LAUNCH((u64)n_polys * B.N, n_batch, k_behz_extend, B, in, in_bstride, n_polys, out, out_bstride);
}
cudaError_t launch_behz_tensor(const BehzC& B, int base, const u64* A, u64 a_bstride, const u64* S, u64 s_bstride, u64* D,
                               u64 d_bstride, u32 n_entries, u32 dim, int s1, int n_batch, cudaStream_t st) {
  if (s1 < 1 || s1 > PIRB_MAX_DIMS + 2 || dim == 0) return cudaErrorInvalidValue;
  const u64 nm = base ? (u64)B.nB + 1 : (u64)B.k;
  CT_LAUNCH((u64)n_entries * nm * B.N, n_batch, k_behz_tensor, B, base, A, a_bstride, S, s_bstride, D, d_bstride, n_entries,
            dim, s1);
}
cudaError_t launch_behz_floor(const BehzC& B, const u64* Dq, u64 dq_bstride, const u64* Db, u64 db_bstride, u64* out,
                              u64 out_bstride, u32 n_polys, int n_batch, cudaStream_t st) {
  CT_LAUNCH((u64)n_polys * B.N, n_batch, k_behz_floor, B, Dq, dq_bstride, Db, db_bstride, out, out_bstride, n_polys);
}
cudaError_t launch_relin_digits(const BehzC& B, const u64* prod, u64 p_bstride, u64* dig, u64 dig_bstride, u32 n_entries,
                                int n_batch, cudaStream_t st) {
  CT_LAUNCH((u64)n_entries * B.k * (B.k + 1) * B.N, n_batch, k_relin_digits, B, prod, p_bstride, dig, dig_bstride, n_entries);
}
cudaError_t launch_relin_mac(const BehzC& B, const u64* dig, u64 dig_bstride, const u64* key, u64* acc, u64 acc_bstride,
                             u32 n_entries, int n_batch, cudaStream_t st) {
  CT_LAUNCH((u64)n_entries * 2 * (B.k + 1) * B.N, n_batch, k_relin_mac, B, dig, dig_bstride, key, acc, acc_bstride, n_entries);
}
cudaError_t launch_relin_finish(const BehzC& B, const u64* prod, u64 p_bstride, const u64* acc, u64 acc_bstride, u64* X,
                                u64 x_bstride, u32 n_entries, int n_batch, cudaStream_t st) {
  CT_LAUNCH((u64)n_entries * 2 * B.k * B.N, n_batch, k_relin_finish, B, prod, p_bstride, acc, acc_bstride, X, x_bstride,
            n_entries);
}
cudaError_t launch_ct_reduce(const BehzC& B, const u64* X, u64 x_bstride, u64* out, u64 out_bstride, u32 n_entries, u32 dim,
                             u32 polys, int n_batch, cudaStream_t st) {
  if (dim == 0 || polys == 0) return cudaErrorInvalidValue;
  const u64 n_groups = (n_entries + dim - 1) / dim;
  CT_LAUNCH(n_groups * polys * B.k * B.N, n_batch, k_ct_reduce, B, X, x_bstride, out, out_bstride, n_entries, dim, polys);
}
#undef CT_LAUNCH

}  // namespace pirb
