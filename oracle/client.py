"""Harness-only restatement of the reference CLIENT side and parameter math (TEST INFRASTRUCTURE ONLY).

Follows pir/cpp/parameters.cpp:56-107 (CreatePIRParameters), pir/cpp/client.cpp:92-144 (createQueryFor),
pir/cpp/client.cpp:196-255 (ProcessReplyCiphertextMult / ProcessReplyCiphertextDecomp), pir/cpp/database.cpp:318-332 (index math),
and SEAL's Plaintext hex-polynomial strings / IntegerEncoder (used by the reference's tests).
"""
from dataclasses import dataclass, field
from typing import List

import numpy as np

from . import binding as ob


@dataclass
class PIRParameters:
    """Mirror of the PIRParameters proto (pir/proto/payload.proto:45-69) with the SEAL params unpacked."""
    num_items: int
    num_pt: int
    dimensions: List[int]
    bytes_per_item: int
    items_per_plaintext: int
    bits_per_coeff: int
    poly_modulus_degree: int
    plain_modulus: int
    coeff_modulus: List[int] = field(default_factory=list)  # data moduli + special prime (last)
    use_ciphertext_multiplication: bool = False

    @property
    def dim_sum(self):
        return sum(self.dimensions)


def create_pir_parameters(dbsize, bytes_per_item=0, dimensions=1, N=4096, plain_bits=20, bits_per_coeff=0,
                          plain_modulus=None, coeff_modulus=None, use_ciphertext_multiplication=False):
    """parameters.cpp:56-107."""
    t = plain_modulus if plain_modulus is not None else ob.plain_modulus_batching(N, plain_bits)
    moduli = coeff_modulus if coeff_modulus is not None else ob.bfv_default(N)
    # StringEncoder ctor, string_encoder.cpp:85: pir::log2 takes a uint32_t, so a wider plain modulus (the reference
    # tests one of 42 bits, correctness_test.cpp:100) is truncated to its low 32 bits first
    enc_bits = ob.log2(t & 0xFFFFFFFF)
    bpc = 0
    if bits_per_coeff > 0:
        if bits_per_coeff > enc_bits:
            raise ValueError("Bits per coefficient greater than max")
        enc_bits = bits_per_coeff
        bpc = bits_per_coeff
    if bytes_per_item > 0:
        items_per_pt = N * enc_bits // bytes_per_item // 8
        if items_per_pt <= 0:
            raise ValueError("Cannot fit an item within one plaintext")
        num_pt = dbsize // items_per_pt
        while dbsize > num_pt * items_per_pt:
            num_pt += 1
        bpi = bytes_per_item
    else:
        bpi = N * enc_bits // 8
        items_per_pt = 1
        num_pt = dbsize
    dims = ob.calculate_dimensions(num_pt, dimensions)
    return PIRParameters(dbsize, num_pt, dims, bpi, items_per_pt, bpc, N, t, list(moduli),
                         bool(use_ciphertext_multiplication))


def calculate_indices(params: PIRParameters, index: int):
    """database.cpp:318-326."""
    pt_index = index // params.items_per_plaintext
    res = [0] * len(params.dimensions)
    for i in range(len(res) - 1, -1, -1):
        res[i] = pt_index % params.dimensions[i]
        pt_index //= params.dimensions[i]
    return res


def calculate_item_offset(params: PIRParameters, index: int):
    """database.cpp:328-332."""
    pt_index = index // params.items_per_plaintext
    return (index - pt_index * params.items_per_plaintext) * params.bytes_per_item


# ---------------------------------------------------------------------------------------------
# SEAL Plaintext hex strings ("4x^4 + FBFCEx^3 + 42") and IntegerEncoder
# ---------------------------------------------------------------------------------------------
def parse_hex_poly(s: str, N: int):
    pt = np.zeros(N, dtype=np.uint64)
    for term in s.replace(" ", "").split("+"):
        if not term:
            continue
        if "x^" in term:
            c, e = term.split("x^")
            pt[int(e)] = int(c, 16)
        else:
            pt[0] = int(term, 16)
    return pt


def format_hex_poly(pt):
    terms = []
    for e in range(len(pt) - 1, -1, -1):
        c = int(pt[e])
        if c:
            terms.append(("%Xx^%d" % (c, e)) if e else ("%X" % c))
    return " + ".join(terms) if terms else "0"


def integer_encode(value: int, N: int, t: int):
    """SEAL IntegerEncoder::encode(int64): binary digits; negative numbers use coefficient t-1."""
    pt = np.zeros(N, dtype=np.uint64)
    neg = value < 0
    v = -value if neg else value
    i = 0
    while v:
        if v & 1:
            pt[i] = (t - 1) if neg else 1
        v >>= 1
        i += 1
    return pt


def integer_decode(pt, t: int):
    """SEAL IntegerEncoder::decode_int64: evaluate at x=2 with centred coefficients."""
    thr = (t + 1) >> 1
    res = 0
    for e in range(len(pt) - 1, -1, -1):
        c = int(pt[e])
        c = c - t if c >= thr else c
        res = 2 * res + c
    return res


# ---------------------------------------------------------------------------------------------
class HarnessClient:
    """PIRClient stand-in: keygen, CreateRequest (raw limbs), ProcessResponse."""

    def __init__(self, params: PIRParameters, seed=1, oracle=None):
        self.params = params
        self.orc = oracle or ob.Oracle(params.poly_modulus_degree, params.coeff_modulus, params.plain_modulus)
        self.keys = self.orc.keygen(seed)
        self.elts = ob.generate_galois_elts(params.poly_modulus_degree)
        self.galois = self.orc.galois_keys(self.keys, self.elts, seed + 1000)
        self._relin = None
        self._seed = seed * 7919 + 17

    @property
    def relin(self):
        """client.cpp:49: keygen_->relin_keys() — one key (for s^2), same layout as one Galois key."""
        if self._relin is None:
            self._relin = self.orc.relin_key(self.keys, self._seed * 31 + 5)
        return self._relin

    def _next_seed(self):
        self._seed += 1
        return self._seed

    def encrypt(self, pt):
        return self.orc.encrypt(self.keys, pt, self._next_seed())

    def decrypt(self, ct, with_budget=False):
        return self.orc.decrypt(self.keys, ct, with_budget)

    def create_query(self, desired_index: int):
        """client.cpp:92-144. Returns [n_ct][2][k][N] uint64."""
        p = self.params
        if desired_index >= p.num_items:
            raise ValueError("invalid index %d" % desired_index)
        N, t = p.poly_modulus_degree, p.plain_modulus
        dims = list(p.dimensions)
        indices = calculate_indices(p, desired_index)
        dim_sum = p.dim_sum
        offset = 0
        n_ct = dim_sum // N + 1
        query = np.zeros((n_ct, 2, self.orc.k, N), dtype=np.uint64)
        for c in range(n_ct):
            pt = np.zeros(N, dtype=np.uint64)
            while indices:
                if indices[0] + offset >= N:
                    indices[0] -= (N - offset)
                    dims[0] -= (N - offset)
                    offset = 0
                    break
                m = N if c < n_ct - 1 else ob.next_power_two(dim_sum % N)
                pt[indices[0] + offset] = pow(m, -1, t)
                offset += dims[0]
                indices.pop(0)
                dims.pop(0)
                if offset >= N:
                    offset -= N
                    break
            query[c] = self.encrypt(pt)
        return query

    def process_reply(self, reply_cts, with_budget=False):
        """client.cpp:219-255: reply [n][2][k][N] -> plaintext coefficients [N]."""
        orc = self.orc
        exp_ratio = orc.ER * 2
        nd = len(self.params.dimensions)
        expect = exp_ratio ** (nd - 1)
        cts = [np.asarray(c) for c in reply_cts]
        if len(cts) != expect:
            raise ValueError("Number of ciphertexts in reply does not match expected")
        min_budget = 1 << 30
        pts = []
        for _ in range(nd):
            pts = []
            for c in cts:
                pt, b = orc.decrypt(self.keys, c, True)
                min_budget = min(min_budget, b)
                pts.append(pt)
            if len(pts) <= 1:
                break
            cts = [orc.reencode_decode(np.stack(pts[i * exp_ratio:(i + 1) * exp_ratio]))
                   for i in range(len(cts) // exp_ratio)]
        return (pts[0], min_budget) if with_budget else pts[0]

    def process_reply_ct(self, reply, with_budget=False):
        """client.cpp:196-217 (ProcessReplyCiphertextMult): the reply is ONE ciphertext ([polys][k][N])."""
        r = np.asarray(reply)
        if r.ndim == 4:
            if r.shape[0] != 1:
                raise ValueError("Number of ciphertexts in reply must be 1 when using CT multiplication")
            r = r[0]
        return self.orc.decrypt_polys(self.keys, r, with_budget)

    def process_response_strings(self, indexes, replies):
        """client.cpp:161-184."""
        p = self.params
        bits = p.bits_per_coeff if p.bits_per_coeff > 0 else ob.log2(p.plain_modulus & 0xFFFFFFFF)
        out = []
        for idx, reply in zip(indexes, replies):
            pt = self.process_reply_ct(reply) if p.use_ciphertext_multiplication else self.process_reply(reply)
            out.append(ob.string_decode(pt, bits, p.bytes_per_item, calculate_item_offset(p, idx)))
        return out


def encode_string_db(params: PIRParameters, items: List[bytes]):
    """PIRDatabase::populate(vector<string>) packing step (database.cpp:84-110) -> [num_pt][N] coefficients."""
    N = params.poly_modulus_degree
    bits = params.bits_per_coeff if params.bits_per_coeff > 0 else ob.log2(params.plain_modulus & 0xFFFFFFFF)
    if len(items) != params.num_items:
        raise ValueError("Database size %d does not match params value %d" % (len(items), params.num_items))
    out = np.zeros((params.num_pt, N), dtype=np.uint64)
    ipp = params.items_per_plaintext
    for i in range(params.num_pt):
        blob = b"".join(items[i * ipp:(i + 1) * ipp])
        c = ob.string_encode(blob, bits, N)
        out[i, :len(c)] = c
    return out


def encode_int_db(params: PIRParameters, values: List[int]):
    """PIRDatabase::populate(vector<int64>) (database.cpp:60-82) -> [num_items][N] coefficients."""
    N, t = params.poly_modulus_degree, params.plain_modulus
    if len(values) != params.num_items:
        raise ValueError("Database size mismatch")
    return np.stack([integer_encode(v, N, t) for v in values])


def db_to_ntt(orc: "ob.Oracle", coeffs):
    """transform_to_ntt_inplace(pt, first_parms_id) for every plaintext -> [num_pt][k][N]."""
    return np.stack([orc.plain_to_ntt(c) for c in coeffs])
