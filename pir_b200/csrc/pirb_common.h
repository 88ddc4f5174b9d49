// pirb_common.h — types shared by host and device code of the B200 PIR answer path.
#pragma once
#include <cstdint>

#include <vector_types.h>  // double2

typedef unsigned long long u64;  // == uint64_t on LP64; matches CUDA's 64-bit intrinsics
typedef uint32_t u32;
typedef uint8_t u8;

#define PIRB_MAX_MODULI 9    // key-level moduli (k data + 1 special); BFVDefault(16384) has 9
#define PIRB_MAX_REENC 64    // 2 * ExpansionRatio upper bound
#define PIRB_MAX_DIMS 8
#define PIRB_MAX_DATA 8      // data-level moduli
#define PIRB_MAX_BSK 10      // auxiliary BEHZ base Bsk = B u {m_sk}: |B| <= |q| + 1

// Per-modulus constants.  Tables live in device memory.
struct ModC {
  u64 q;
  u64 ratio_hi, ratio_lo;  // floor(2^128 / q)
  u64 inv_n, inv_n_s;      // N^{-1} mod q and its Shoup companion floor(inv_n * 2^64 / q)
  const u64* rp;           // rp[bitrev(i)] = psi^i  (psi = minimal primitive 2N-th root; SEAL ordering)
  const u64* rps;          // Shoup companions
  const u64* irp;          // irp[i] = rp[i]^{-1}
  const u64* irps;
  // FP64 engine (moduli <= 44 bits): the same constants as integer-valued doubles, plus x/q companions
  double qd, qinv;         // q and 1/q
  // each table entry is the pair (w, w/q) so one 128-bit load fetches both operands of f64_modmul
  const double2* fw;       // fw[i].x  = (double)rp[i]                    forward twiddles, SEAL ordering
  const double2* iw;       // iw[g + j].x = psi^(-j*N/g)  for gap g = 1,2,4,...,N/2, j < g   (inverse, DIT form)
  const double2* fin;      // fin[i].x = N^{-1} * psi^{-i}                final scaling of the inverse
  const double* fw1;       // w-only copies of fw / iw (N doubles each): bulk-copied into shared memory by the
  const double* iw1;       // cluster key-switch kernel, which forms w/q on the fly
  double pow_h, pow_h_i;    // 2^h mod q and (2^h mod q)/q     (h = DevParams::half_bits; FP64 Karatsuba recombination)
  double pow_2h, pow_2h_i;  // 2^(2h) mod q and its /q companion
};

// Passed by value (__grid_constant__) to every kernel: lives in the constant bank.
struct DevParams {
  int logn;
  int k;        // data-level moduli; m[k] is the special prime P
  u32 N;
  u32 ptb;      // trunc(log2 t)
  u64 t;
  u64 thr;      // (t+1)/2  plain_upper_half_threshold
  u64 half_P;   // P >> 1
  ModC m[PIRB_MAX_MODULI];
  u64 inv_P[PIRB_MAX_MODULI];       // P^{-1} mod q_j
  u64 inv_P_s[PIRB_MAX_MODULI];     // Shoup companion
  u64 half_P_mod[PIRB_MAX_MODULI];  // (P>>1) mod q_j
  double inv_P_d[PIRB_MAX_MODULI];  // the same constants as doubles for the FP64 engine's mod-down
  double inv_P_di[PIRB_MAX_MODULI]; // inv_P / q_j
  double half_P_mod_d[PIRB_MAX_MODULI];
  double half_P_d;
  int two_er;                       // 2 * ExpansionRatio
  int lazy_ntt;                     // 1 if every modulus is below 2^(62 - log2 N): fully lazy butterflies
  int ntt_engine;                   // 0 integer, 1 integer lazy, 2 FP64 (moduli <= 44 bits)
  int mac_mode;                     // lazy MAC flavour the moduli allow: 0 wide (any), 1 int24 (< 2^48), 2 fp64 (<= 44 bit)
  int half_bits;                    // h: operand split position for the fp64 MAC (ceil(max modulus bits / 2))
  u32 mac_max_terms;                // longest exact accumulation chain for mac_mode
  u32 wide_max_terms;               // ... and for the 128-bit accumulator: floor(2^128 / q_max^2), capped at 2^30
  u8 re_poly[PIRB_MAX_REENC];       // re-encode chunk e -> source poly (0/1)
  u8 re_mod[PIRB_MAX_REENC];        //                  -> source modulus j
  u8 re_shift[PIRB_MAX_REENC];      //                  -> right shift
};


// ---- ciphertext-multiplication mode (database.cpp:202-211): SEAL 3.5.6 RNSTool constants for the first data level ----
struct ModLite {
  u64 q;
  u64 ratio_hi, ratio_lo;  // floor(2^128 / q)
};
// Bases: q = data moduli, B = nB auxiliary 61-bit primes, Bsk = B u {m_sk} (m_sk last), m_tilde = 2^32.
// Passed by value to the kernels of kernels_ctmul.cu (constant bank).
struct BehzC {
  int k, nB;
  u32 N;
  int logn;
  ModLite q[PIRB_MAX_DATA];
  ModLite bsk[PIRB_MAX_BSK];
  u64 t_mod_q[PIRB_MAX_DATA], t_mod_bsk[PIRB_MAX_BSK];
  u64 mtilde_mod_q[PIRB_MAX_DATA];                 // 2^32 mod q_j
  u64 inv_qhat_mod_q[PIRB_MAX_DATA];               // (Q/q_j)^-1 mod q_j
  u64 qhat_mod_bsk[PIRB_MAX_BSK][PIRB_MAX_DATA];   // (Q/q_j) mod bsk_i
  u64 qhat_mod_mtilde[PIRB_MAX_DATA];              // (Q/q_j) mod 2^32
  u64 neg_inv_q_mod_mtilde;                        // -Q^-1 mod 2^32
  u64 q_mod_bsk[PIRB_MAX_BSK];                     // Q mod bsk_i
  u64 inv_mtilde_mod_bsk[PIRB_MAX_BSK];            // 2^-32 mod bsk_i
  u64 inv_q_mod_bsk[PIRB_MAX_BSK];                 // Q^-1 mod bsk_i
  u64 inv_bhat_mod_b[PIRB_MAX_BSK];                // (B/b_j)^-1 mod b_j
  u64 bhat_mod_q[PIRB_MAX_DATA][PIRB_MAX_BSK];     // (B/b_j) mod q_i
  u64 bhat_mod_msk[PIRB_MAX_BSK];                  // (B/b_j) mod m_sk
  u64 inv_b_mod_msk;                               // B^-1 mod m_sk
  u64 b_mod_q[PIRB_MAX_DATA];                      // B mod q_i
  // relinearization (switch_key_inplace): the special prime and its mod-down constants
  ModLite P;
  u64 half_P;
  u64 half_P_mod_q[PIRB_MAX_DATA], inv_P_mod_q[PIRB_MAX_DATA];
};
static_assert(sizeof(BehzC) <= 4000, "BehzC must fit the kernel parameter space");
