"""ctypes binding to the CPU oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (pir_b200/) never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)
i8p = C.POINTER(C.c_int8)


def build(force=False):
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_uint64, C.c_uint32, u64p, C.c_uint64]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_expansion_ratio.restype = C.c_uint32
        L.orc_expansion_ratio.argtypes = [C.c_void_p]
        L.orc_psi.restype = C.c_uint64
        L.orc_psi.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_plain_modulus_batching.restype = C.c_uint64
        L.orc_plain_modulus_batching.argtypes = [C.c_uint64, C.c_int]
        L.orc_bfv_default.argtypes = [C.c_uint64, u64p, C.c_uint32]
        L.orc_is_prime.argtypes = [C.c_uint64]
        L.orc_calculate_dimensions.argtypes = [C.c_uint32, C.c_uint32, u32p]
        L.orc_next_power_two.restype = C.c_uint64
        L.orc_next_power_two.argtypes = [C.c_uint64]
        L.orc_ceil_log2.restype = C.c_uint32
        L.orc_ceil_log2.argtypes = [C.c_uint32]
        L.orc_log2.restype = C.c_uint32
        L.orc_log2.argtypes = [C.c_uint32]
        L.orc_ntt_forward.argtypes = [C.c_void_p, C.c_uint32, u64p]
        L.orc_ntt_inverse.argtypes = [C.c_void_p, C.c_uint32, u64p]
        L.orc_mulmod.restype = C.c_uint64
        L.orc_mulmod.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64]
        L.orc_plain_to_ntt.argtypes = [C.c_void_p, u64p, C.c_uint64, u64p]
        L.orc_ct_to_ntt.argtypes = [C.c_void_p, u64p]
        L.orc_ct_from_ntt.argtypes = [C.c_void_p, u64p]
        L.orc_substitute.argtypes = [C.c_void_p, u64p, C.c_uint32, u32p, C.c_uint32, u64p]
        L.orc_mul_inv_pow_x.argtypes = [C.c_void_p, u64p, C.c_uint32, u64p]
        L.orc_expand.argtypes = [C.c_void_p, u64p, C.c_uint64, C.c_uint64, u32p, C.c_uint32, u64p, u64p, C.c_int]
        L.orc_reencode.argtypes = [C.c_void_p, u64p, u64p]
        L.orc_reencode_decode.argtypes = [C.c_void_p, u64p, u64p]
        L.orc_db_multiply.argtypes = [C.c_void_p, u64p, C.c_uint64, u32p, C.c_uint32, u64p, C.c_uint64, u64p,
                                      C.c_uint64, u64p]
        L.orc_process_query.argtypes = [C.c_void_p, u64p, C.c_uint64, u32p, C.c_uint32, u32p, C.c_uint32, u64p, u64p,
                                        C.c_uint64, u64p, C.c_uint64, u64p]
        L.orc_scan_row.argtypes = [C.c_void_p, u64p, C.c_uint64, u64p, u64p]
        L.orc_string_encode.restype = C.c_int64
        L.orc_string_encode.argtypes = [u8p, C.c_uint64, C.c_uint64, u64p, C.c_uint64]
        L.orc_string_decode.argtypes = [u64p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, u8p]
        L.orc_keygen.argtypes = [C.c_void_p, C.c_uint64, u64p, i8p, u64p]
        L.orc_gen_galois_keys.argtypes = [C.c_void_p, C.c_uint64, u64p, i8p, u32p, C.c_uint32, u64p]
        L.orc_encrypt.argtypes = [C.c_void_p, C.c_uint64, u64p, u64p, C.c_uint64, u64p]
        L.orc_decrypt.argtypes = [C.c_void_p, u64p, u64p, u64p]
        L.orc_rns_bases.restype = C.c_uint32
        L.orc_rns_bases.argtypes = [C.c_void_p, u64p, C.c_uint32]
        L.orc_bfv_multiply.argtypes = [C.c_void_p, u64p, C.c_uint32, u64p, C.c_uint32, u64p]
        L.orc_relinearize.argtypes = [C.c_void_p, u64p, u64p]
        L.orc_gen_relin_key.argtypes = [C.c_void_p, C.c_uint64, u64p, i8p, u64p]
        L.orc_decrypt_polys.argtypes = [C.c_void_p, u64p, u64p, C.c_uint32, u64p]
        L.orc_db_multiply_ct.argtypes = [C.c_void_p, u64p, C.c_uint64, u32p, C.c_uint32, u64p, C.c_uint64, u64p, u64p,
                                         C.c_uint32, u32p]
        L.orc_process_query_ct.argtypes = [C.c_void_p, u64p, C.c_uint64, u32p, C.c_uint32, u32p, C.c_uint32, u64p, u64p,
                                           u64p, C.c_uint64, u64p, C.c_uint32, u32p]
        _lib = L
    return _lib


def _p64(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


def _p32(a):
    assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u32p)


# ---------------------------------------------------------------------------------------------
# parameter math (module-level)
# ---------------------------------------------------------------------------------------------
def plain_modulus_batching(N, bits):
    return int(lib().orc_plain_modulus_batching(N, bits))


def bfv_default(N):
    buf = np.zeros(16, dtype=np.uint64)
    n = lib().orc_bfv_default(N, _p64(buf), 16)
    if n < 0:
        raise ValueError("no BFVDefault for N=%d" % N)
    return [int(x) for x in buf[:n]]


def is_prime(v):
    return bool(lib().orc_is_prime(v))


def calculate_dimensions(db_size, nd):
    out = np.zeros(nd, dtype=np.uint32)
    lib().orc_calculate_dimensions(db_size, nd, _p32(out))
    return [int(x) for x in out]


def next_power_two(v):
    return int(lib().orc_next_power_two(v))


def ceil_log2(v):
    return int(lib().orc_ceil_log2(v))


def log2(v):
    return int(lib().orc_log2(v))


def generate_galois_elts(N):
    return [(N >> i) + 1 for i in range(ceil_log2(N))]


def string_encode(data: bytes, bits_per_coeff: int, max_coeff: int):
    buf = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data if data else b"\0")
    out = np.zeros(max_coeff, dtype=np.uint64)
    n = lib().orc_string_encode(buf, len(data), bits_per_coeff, _p64(out), max_coeff)
    if n < 0:
        raise ValueError("Number of coefficients needed greater than poly modulus degree")
    return out[:n].copy()


def string_decode(coeffs, bits_per_coeff, length, byte_offset=0):
    coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64)
    out = (C.c_uint8 * max(1, length))()
    rc = lib().orc_string_decode(_p64(coeffs), len(coeffs), bits_per_coeff, length, byte_offset, out)
    if rc:
        raise ValueError("Requested decode beyond end of data in polynomial")
    return bytes(out[:length])


# ---------------------------------------------------------------------------------------------
class Oracle:
    """One BFV parameter set: N, data moduli q[0..k), special prime P (last of `moduli`), plain modulus t."""

    def __init__(self, N, moduli, t):
        self.N = int(N)
        self.moduli = [int(m) for m in moduli]
        self.k = len(self.moduli) - 1
        self.t = int(t)
        m = np.array(self.moduli, dtype=np.uint64)
        self.h = lib().orc_create(self.N, len(self.moduli), _p64(m), self.t)
        if not self.h:
            raise ValueError("invalid oracle parameters")
        self.ct_limbs = 2 * self.k * self.N
        self.pt_limbs = self.k * self.N
        self.key_limbs = self.k * 2 * (self.k + 1) * self.N
        self.ER = int(lib().orc_expansion_ratio(self.h))
        self.ptb = log2(self.t)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    @classmethod
    def default(cls, N=4096, plain_bits=20):
        return cls(N, bfv_default(N), plain_modulus_batching(N, plain_bits))

    def psi(self, j):
        return int(lib().orc_psi(self.h, j))

    # ring primitives ---------------------------------------------------------------------
    def ntt_forward(self, j, poly):
        a = np.ascontiguousarray(poly, dtype=np.uint64).copy()
        lib().orc_ntt_forward(self.h, j, _p64(a))
        return a

    def ntt_inverse(self, j, poly):
        a = np.ascontiguousarray(poly, dtype=np.uint64).copy()
        lib().orc_ntt_inverse(self.h, j, _p64(a))
        return a

    def plain_to_ntt(self, coeffs):
        c = np.ascontiguousarray(coeffs, dtype=np.uint64)
        out = np.zeros(self.pt_limbs, dtype=np.uint64)
        lib().orc_plain_to_ntt(self.h, _p64(c), len(c), _p64(out))
        return out.reshape(self.k, self.N)

    def ct_to_ntt(self, ct):
        a = np.ascontiguousarray(ct, dtype=np.uint64).copy()
        lib().orc_ct_to_ntt(self.h, _p64(a))
        return a

    def ct_from_ntt(self, ct):
        a = np.ascontiguousarray(ct, dtype=np.uint64).copy()
        lib().orc_ct_from_ntt(self.h, _p64(a))
        return a

    # server path -------------------------------------------------------------------------
    def substitute(self, ct, power, elts, keys):
        a = np.ascontiguousarray(ct, dtype=np.uint64).copy()
        e = np.array(elts, dtype=np.uint32)
        rc = lib().orc_substitute(self.h, _p64(a), power, _p32(e), len(e), _p64(keys))
        if rc:
            raise RuntimeError("oracle substitute rc=%d" % rc)
        return a

    def mul_inv_pow_x(self, ct, k):
        a = np.ascontiguousarray(ct, dtype=np.uint64)
        out = np.zeros_like(a)
        lib().orc_mul_inv_pow_x(self.h, _p64(a), k, _p64(out))
        return out

    def expand(self, cts, total_items, elts, keys, single=False):
        a = np.ascontiguousarray(cts, dtype=np.uint64).reshape(-1, self.ct_limbs)
        e = np.array(elts, dtype=np.uint32)
        out = np.zeros((max(1, total_items), self.ct_limbs), dtype=np.uint64)
        rc = lib().orc_expand(self.h, _p64(a), a.shape[0], total_items, _p32(e), len(e), _p64(keys), _p64(out),
                              1 if single else 0)
        if rc:
            raise OracleStatus(rc)
        return out[:total_items].reshape(total_items, 2, self.k, self.N)

    def reencode(self, ct):
        a = np.ascontiguousarray(ct, dtype=np.uint64)
        out = np.zeros((2 * self.ER, self.N), dtype=np.uint64)
        lib().orc_reencode(self.h, _p64(a), _p64(out))
        return out

    def reencode_decode(self, pts):
        a = np.ascontiguousarray(pts, dtype=np.uint64)
        out = np.zeros((2, self.k, self.N), dtype=np.uint64)
        lib().orc_reencode_decode(self.h, _p64(a), _p64(out))
        return out

    def db_multiply(self, db_ntt, dims, sv):
        """db_ntt [num_pt][k][N]; sv [n_sv][2][k][N] (mutated: returned too). Returns (reply cts, sv_after)."""
        db = np.ascontiguousarray(db_ntt, dtype=np.uint64)
        num_pt = db.size // self.pt_limbs
        svc = np.ascontiguousarray(sv, dtype=np.uint64).copy()
        n_sv = svc.size // self.ct_limbs
        d = np.array(dims, dtype=np.uint32)
        cap = (2 * self.ER) ** (len(dims) - 1)
        out = np.zeros((cap, 2, self.k, self.N), dtype=np.uint64)
        cnt = C.c_uint64(0)
        rc = lib().orc_db_multiply(self.h, _p64(db), num_pt, _p32(d), len(d), _p64(svc), n_sv, _p64(out), cap,
                                   C.byref(cnt))
        if rc:
            raise OracleStatus(rc)
        return out[:cnt.value], svc

    def process_query(self, db_ntt, dims, elts, keys, query):
        db = np.ascontiguousarray(db_ntt, dtype=np.uint64)
        num_pt = db.size // self.pt_limbs
        q = np.ascontiguousarray(query, dtype=np.uint64)
        n_ct = q.size // self.ct_limbs
        d = np.array(dims, dtype=np.uint32)
        e = np.array(elts, dtype=np.uint32)
        cap = (2 * self.ER) ** (len(dims) - 1)
        out = np.zeros((cap, 2, self.k, self.N), dtype=np.uint64)
        cnt = C.c_uint64(0)
        rc = lib().orc_process_query(self.h, _p64(db), num_pt, _p32(d), len(d), _p32(e), len(e), _p64(keys), _p64(q),
                                     n_ct, _p64(out), cap, C.byref(cnt))
        if rc:
            raise OracleStatus(rc)
        return out[:cnt.value]

    # ciphertext-multiplication mode (database.cpp:202-211; oracle/bfv_mul_oracle.hpp) ---------
    def rns_bases(self):
        """(m_sk, [primes of B]) — the auxiliary BEHZ bases SEAL's RNSTool picks for the first data level."""
        buf = np.zeros(16, dtype=np.uint64)
        nb = int(lib().orc_rns_bases(self.h, _p64(buf), 16))
        return int(buf[0]), [int(x) for x in buf[1:1 + nb]]

    def bfv_multiply(self, a, b):
        """a [s1][k][N] x b [s2][k][N] (coefficient form) -> [s1+s2-1][k][N]  (Evaluator::multiply)."""
        a = np.ascontiguousarray(a, dtype=np.uint64)
        b = np.ascontiguousarray(b, dtype=np.uint64)
        s1, s2 = a.size // self.pt_limbs, b.size // self.pt_limbs
        out = np.zeros((s1 + s2 - 1, self.k, self.N), dtype=np.uint64)
        lib().orc_bfv_multiply(self.h, _p64(a), s1, _p64(b), s2, _p64(out))
        return out

    def relinearize(self, ct3, relin_key):
        """[3][k][N] -> [2][k][N]  (Evaluator::relinearize_inplace with one relinearization key)."""
        a = np.ascontiguousarray(ct3, dtype=np.uint64).copy().reshape(3, self.k, self.N)
        lib().orc_relinearize(self.h, _p64(a), _p64(relin_key))
        return a[:2].copy()

    def relin_key(self, keys, seed):
        out = np.zeros(self.key_limbs, dtype=np.uint64)
        lib().orc_gen_relin_key(self.h, seed, _p64(keys["sk_ntt"]), keys["sk_coeff"].ctypes.data_as(i8p), _p64(out))
        return out

    def decrypt_polys(self, keys, ct, with_budget=False):
        a = np.ascontiguousarray(ct, dtype=np.uint64)
        polys = a.size // self.pt_limbs
        pt = np.zeros(self.N, dtype=np.uint64)
        budget = lib().orc_decrypt_polys(self.h, _p64(keys["sk_ntt"]), _p64(a), polys, _p64(pt))
        return (pt, budget) if with_budget else pt

    def db_multiply_ct(self, db_ntt, dims, sv, relin_key=None):
        """PIRDatabase::multiply with use_ciphertext_multiplication: one ciphertext [polys][k][N] (polys = 2 with a
        relinearization key, else one more per upper dimension).  The selection vector is not modified."""
        db = np.ascontiguousarray(db_ntt, dtype=np.uint64)
        num_pt = db.size // self.pt_limbs
        svc = np.ascontiguousarray(sv, dtype=np.uint64)
        n_sv = svc.size // self.ct_limbs
        d = np.array(dims, dtype=np.uint32)
        cap = len(dims) + 1
        out = np.zeros((cap, self.k, self.N), dtype=np.uint64)
        polys = C.c_uint32(0)
        rk = _p64(relin_key) if relin_key is not None else None
        rc = lib().orc_db_multiply_ct(self.h, _p64(db), num_pt, _p32(d), len(d), _p64(svc), n_sv, rk, _p64(out), cap,
                                      C.byref(polys))
        if rc:
            raise OracleStatus(rc)
        return out[:polys.value].copy()

    def process_query_ct(self, db_ntt, dims, elts, keys, query, relin_key=None):
        db = np.ascontiguousarray(db_ntt, dtype=np.uint64)
        num_pt = db.size // self.pt_limbs
        q = np.ascontiguousarray(query, dtype=np.uint64)
        n_ct = q.size // self.ct_limbs
        d = np.array(dims, dtype=np.uint32)
        e = np.array(elts, dtype=np.uint32)
        cap = len(dims) + 1
        out = np.zeros((cap, self.k, self.N), dtype=np.uint64)
        polys = C.c_uint32(0)
        rk = _p64(relin_key) if relin_key is not None else None
        rc = lib().orc_process_query_ct(self.h, _p64(db), num_pt, _p32(d), len(d), _p32(e), len(e), _p64(keys), rk,
                                        _p64(q), n_ct, _p64(out), cap, C.byref(polys))
        if rc:
            raise OracleStatus(rc)
        return out[:polys.value].copy()

    def scan_row(self, db_ntt, sv_ntt):
        db = np.ascontiguousarray(db_ntt, dtype=np.uint64)
        sv = np.ascontiguousarray(sv_ntt, dtype=np.uint64)
        count = db.size // self.pt_limbs
        out = np.zeros((2, self.k, self.N), dtype=np.uint64)
        lib().orc_scan_row(self.h, _p64(db), count, _p64(sv), _p64(out))
        return out

    # harness crypto ----------------------------------------------------------------------
    def keygen(self, seed):
        sk_ntt = np.zeros((self.k + 1) * self.N, dtype=np.uint64)
        sk_coeff = np.zeros(self.N, dtype=np.int8)
        pk = np.zeros(2 * (self.k + 1) * self.N, dtype=np.uint64)
        lib().orc_keygen(self.h, seed, _p64(sk_ntt), sk_coeff.ctypes.data_as(i8p), _p64(pk))
        return {"sk_ntt": sk_ntt, "sk_coeff": sk_coeff, "pk": pk}

    def galois_keys(self, keys, elts, seed):
        e = np.array(elts, dtype=np.uint32)
        out = np.zeros(len(e) * self.key_limbs, dtype=np.uint64)
        lib().orc_gen_galois_keys(self.h, seed, _p64(keys["sk_ntt"]), keys["sk_coeff"].ctypes.data_as(i8p), _p32(e),
                                  len(e), _p64(out))
        return out

    def encrypt(self, keys, pt, seed):
        p = np.ascontiguousarray(pt, dtype=np.uint64)
        ct = np.zeros((2, self.k, self.N), dtype=np.uint64)
        lib().orc_encrypt(self.h, seed, _p64(keys["pk"]), _p64(p), len(p), _p64(ct))
        return ct

    def decrypt(self, keys, ct, with_budget=False):
        a = np.ascontiguousarray(ct, dtype=np.uint64)
        pt = np.zeros(self.N, dtype=np.uint64)
        budget = lib().orc_decrypt(self.h, _p64(keys["sk_ntt"]), _p64(a), _p64(pt))
        return (pt, budget) if with_budget else pt


class OracleStatus(Exception):
    """Mirrors absl::Status codes of the reference: 3 = InvalidArgument, 13 = Internal."""

    def __init__(self, code):
        super().__init__("oracle status %d" % code)
        self.code = code
