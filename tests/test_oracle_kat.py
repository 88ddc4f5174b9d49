"""Pins the CPU oracle against the reference's own known-answer vectors (decrypt level) — CPU only.

Every vector below is transcribed from a reference test (file:line cited, relative to /root/reference).
The reference checks these through decryption with fresh random keys; so do we.
"""
import json
import os

import numpy as np
import pytest

from oracle import binding as ob
from oracle import client as oc

N = 4096
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_kats.json")) as _f:
    GOLD = json.load(_f)  # the reference's own known-answer vectors, with the file:line each set comes from


# ------------------------------------------------------------------ parameters / primes
def test_plain_modulus_batching_pinned():
    # server_test.cpp:295 pins t = 0xFC001 for (4096, 20 bit): "1x^1" -> "FC000x^1"
    assert ob.plain_modulus_batching(4096, 20) == 0xFC001
    assert ob.plain_modulus_batching(4096, 16) == 40961
    assert ob.plain_modulus_batching(4096, 24) == 0xFFC001
    assert ob.plain_modulus_batching(8192, 20) == 0xFC001
    for n, b in [(4096, 20), (4096, 16), (4096, 22), (4096, 24), (8192, 20), (8192, 42)]:
        p = ob.plain_modulus_batching(n, b)
        assert ob.is_prime(p) and p % (2 * n) == 1 and p.bit_length() == b


@pytest.mark.parametrize("n,total_bits", [(4096, 109), (8192, 218), (16384, 438)])
def test_bfv_default_moduli(n, total_bits):
    mods = ob.bfv_default(n)
    assert sum(m.bit_length() for m in mods) == total_bits  # SEAL's 128-bit security budget
    for m in mods:
        assert ob.is_prime(m) and m % (2 * n) == 1
    assert len(set(mods)) == len(mods)


def test_utils_tables():
    # utils_test.cpp:24-63
    for v, e in [(0, 1), (1, 1), (2, 2), (3, 4), (8, 8), (9, 16), (1 << 16, 65536), ((1 << 16) + 1, 131072),
                 ((1 << 30) + 1, 2147483648)]:
        assert ob.next_power_two(v) == e
    for v, e in [(1, 0), (2, 1), (3, 2), (8, 3), (15, 4), (16, 4), (17, 5), ((1 << 16) - 1, 16), (1 << 16, 16),
                 ((1 << 16) + 1, 17), (1 << 31, 31)]:
        assert ob.ceil_log2(v) == e
    for v, e in [(1, 0), (2, 1), (3, 1), (8, 3), (15, 3), (16, 4), (17, 4), ((1 << 16) - 1, 15), (1 << 16, 16),
                 ((1 << 16) + 1, 16), ((1 << 31) - 1, 30), (1 << 31, 31)]:
        assert ob.log2(v) == e
    assert ob.generate_galois_elts(4096) == [4097, 2049, 1025, 513, 257, 129, 65, 33, 17, 9, 5, 3]


def test_calculate_dimensions():
    # database_test.cpp:456-464
    for n, d, e in [(100, 1, [100]), (100, 2, [10, 10]), (82, 2, [10, 9]), (975, 2, [32, 31]),
                    (1000, 3, [10, 10, 10]), (1001, 3, [11, 10, 10]), (1000001, 3, [101, 100, 100])]:
        assert ob.calculate_dimensions(n, d) == e


def test_calculate_indices_and_offsets():
    # database_test.cpp:409-421 (params: N=4096, 16-bit t)
    for items, size, d, idx, exp in [(100, 0, 1, 42, [42]), (100, 0, 1, 7, [7]), (84, 0, 2, 7, [0, 7]),
                                     (87, 0, 2, 27, [3, 0]), (87, 0, 2, 42, [4, 6]), (87, 0, 2, 86, [9, 5]),
                                     (82, 0, 3, 3, [0, 0, 3]), (82, 0, 3, 20, [1, 0, 0]), (82, 0, 3, 75, [3, 3, 3]),
                                     (5000, 64, 1, 2222, [18]), (5000, 64, 1, 1200, [10])]:
        p = oc.create_pir_parameters(items, size, d, 4096, 16)
        assert oc.calculate_indices(p, idx) == exp
    # database_test.cpp:441-444
    for items, size, idx, exp in [(100, 0, 42, 0), (1000, 64, 42, 2688), (1000, 64, 960, 0), (1000, 64, 999, 2496)]:
        p = oc.create_pir_parameters(items, size, 1, 4096, 16)
        assert oc.calculate_item_offset(p, idx) == exp


def test_pir_parameters():
    # parameters_test.cpp:49-98
    p = oc.create_pir_parameters(1026, 256)
    assert (p.num_pt, p.items_per_plaintext, p.dimensions) == (27, 38, [27])
    p = oc.create_pir_parameters(19011, 500, 3)
    assert (p.num_pt, p.items_per_plaintext, p.dimensions) == (1001, 19, [11, 10, 10])
    p = oc.create_pir_parameters(77412, 777, 2, N=8192, bits_per_coeff=12)
    assert (p.num_pt, p.items_per_plaintext, p.dimensions, p.bits_per_coeff) == (5161, 15, [72, 72], 12)


def test_string_encoder_shapes_and_roundtrip():
    # string_encoder_test.cpp:64-71 (N=4096, 20-bit t -> 19 bits per coeff)
    bits = ob.log2(ob.plain_modulus_batching(4096, 20))
    assert bits == 19
    per_pt = lambda sz: 4096 * bits // sz // 8
    assert [per_pt(s) for s in (1, 9728, 9729, 99999, 64, 288)] == [9728, 1, 0, 0, 152, 33]
    # string_encoder_test.cpp:209-211 max bytes
    for n, b, exp in [(4096, 20, 9728), (4096, 16, 7680), (8192, 20, 19456)]:
        assert n * ob.log2(ob.plain_modulus_batching(n, b)) // 8 == exp
    # string_encoder_test.cpp:73-83 encode/decode
    value = b"This is a string test for random VALUES@!#"
    c = ob.string_encode(value, 19, 4096)
    assert len(c) == -(-len(value) * 8 // 19)
    assert ob.string_decode(c, 19, len(value)) == value
    rng = np.random.default_rng(3)
    blob = rng.integers(0, 256, 9728, dtype=np.uint8).tobytes()
    c = ob.string_encode(blob, 19, 4096)
    assert len(c) == 4096 and int(c.max()) < (1 << 19)
    assert ob.string_decode(c, 19, 9728) == blob
    assert ob.string_decode(c, 19, 64, 64 * 77) == blob[64 * 77:64 * 78]
    with pytest.raises(ValueError):
        ob.string_encode(blob + b"x", 19, 4096)  # string_encoder_test: too big
    with pytest.raises(ValueError):
        ob.string_decode(c, 19, 100, 9700)


# ------------------------------------------------------------------ NTT definition
def test_ntt_matches_definition_small():
    # N=16 toy ring with a real NTT-friendly prime; out[i] = a(psi^(2*bitrev(i)+1)), psi minimal.
    n = 16
    q = 0xffffee001  # ≡ 1 mod 8192, so also ≡ 1 mod 32
    o = ob.Oracle(n, [q, 0xffffc4001], 17)
    psi = o.psi(0)
    assert pow(psi, n, q) == q - 1
    # minimality: no smaller primitive 2n-th root exists
    roots = sorted(pow(psi, 2 * i + 1, q) for i in range(n))
    assert roots[0] == psi
    rng = np.random.default_rng(0)
    a = rng.integers(0, q, n, dtype=np.uint64)
    out = o.ntt_forward(0, a)
    br = lambda i: int(format(i, "04b")[::-1], 2)
    for i in range(n):
        x = pow(psi, 2 * br(i) + 1, q)
        assert int(out[i]) == sum(int(a[j]) * pow(x, j, q) for j in range(n)) % q
    assert np.array_equal(o.ntt_inverse(0, out), a)


def test_ntt_roundtrip_and_negacyclic_product(orc4096):
    o = orc4096
    rng = np.random.default_rng(1)
    for j in range(o.k + 1):
        q = o.moduli[j]
        a = rng.integers(0, q, N, dtype=np.uint64)
        assert np.array_equal(o.ntt_inverse(j, o.ntt_forward(j, a)), a)
        # x^(N-1) * x = -1
        x1 = np.zeros(N, dtype=np.uint64); x1[1] = 1
        xl = np.zeros(N, dtype=np.uint64); xl[N - 1] = 1
        f1, fl = o.ntt_forward(j, x1), o.ntt_forward(j, xl)
        prod = np.array([int(u) * int(v) % q for u, v in zip(f1, fl)], dtype=np.uint64)
        res = o.ntt_inverse(j, prod)
        assert int(res[0]) == q - 1 and not res[1:].any()


def test_barrett_mulmod_matches_python(orc4096):
    rng = np.random.default_rng(2)
    for j in range(3):
        q = orc4096.moduli[j]
        for _ in range(200):
            a, b = int(rng.integers(0, q)), int(rng.integers(0, q))
            assert ob.lib().orc_mulmod(orc4096.h, j, a, b) == a * b % q


# ------------------------------------------------------------------ encrypt/decrypt sanity
def test_encrypt_decrypt_roundtrip(client4096):
    c = client4096
    rng = np.random.default_rng(5)
    pt = rng.integers(0, c.orc.t, N, dtype=np.uint64)
    ct = c.encrypt(pt)
    out, budget = c.decrypt(ct, with_budget=True)
    assert np.array_equal(out, pt)
    assert budget > 40  # fresh BFV ct at N=4096 / 72-bit q / 20-bit t


# ------------------------------------------------------------------ server_test.cpp:291-305 (11 vectors)
SUBST = [tuple(v) for v in GOLD["substitute_power_x_inplace"]["vectors"]]


@pytest.mark.parametrize("inp,power,expected", SUBST)
def test_substitute_kat(client4096, inp, power, expected):
    c = client4096
    ct = c.encrypt(oc.parse_hex_poly(inp, N))
    gk = c.orc.galois_keys(c.keys, [power], 77)
    out = c.orc.substitute(ct, power, [power], gk)
    assert oc.format_hex_poly(c.decrypt(out)) == oc.format_hex_poly(oc.parse_hex_poly(expected, N))


def test_substitute_missing_key_is_internal_error(client4096):
    c = client4096
    ct = c.encrypt(oc.parse_hex_poly("1x^1", N))
    gk = c.orc.galois_keys(c.keys, [5], 77)
    with pytest.raises(RuntimeError):
        c.orc.substitute(ct, 3, [5], gk)


# ------------------------------------------------------------------ server_test.cpp:333-339 (4 vectors)
SHIFT = [tuple(v) for v in GOLD["multiply_inverse_power_of_x"]["vectors"]]


@pytest.mark.parametrize("inp,k,expected", SHIFT)
def test_multiply_inverse_power_x_kat(client4096, inp, k, expected):
    c = client4096
    ct = c.encrypt(oc.parse_hex_poly(inp, N))
    out = c.orc.mul_inv_pow_x(ct, k)
    assert oc.format_hex_poly(c.decrypt(out)) == oc.format_hex_poly(oc.parse_hex_poly(expected, N))


# ------------------------------------------------------------------ server_test.cpp:376-383 (4 vectors)
EXPAND = [tuple(v) for v in GOLD["oblivious_expansion"]["vectors"]]


@pytest.mark.parametrize("inp,expected", EXPAND)
def test_oblivious_expansion_kat(client4096, inp, expected):
    c = client4096
    ct = c.encrypt(oc.parse_hex_poly(inp, N))
    out = c.orc.expand(ct, len(expected), c.elts, c.galois, single=True)
    assert len(out) == len(expected)
    for o, e in zip(out, expected):
        assert oc.format_hex_poly(c.decrypt(o)) == oc.format_hex_poly(oc.parse_hex_poly(e, N))


# ------------------------------------------------------------------ server_test.cpp:423-428 (6 cases)
@pytest.mark.parametrize("num_items,index,expected_value",
                         [tuple(v) for v in GOLD["oblivious_expansion_multi_ct"]["vectors"]])
def test_oblivious_expansion_multi_ct(client4096, num_items, index, expected_value):
    c = client4096
    n_ct = num_items // N + 1
    cts = []
    for i in range(n_ct):
        pt = np.zeros(N, dtype=np.uint64)
        if index // N == i:
            pt[index % N] = 1
        cts.append(c.encrypt(pt))
    out = c.orc.expand(np.stack(cts), num_items, c.elts, c.galois)
    assert len(out) == num_items
    # decrypting thousands of cts is slow in the harness; check the hot one, its neighbours and a stride
    probe = sorted(set([index, 0, num_items - 1, max(0, index - 1), min(num_items - 1, index + 1)] +
                       list(range(0, num_items, 257))))
    for i in probe:
        pt = c.decrypt(out[i])
        assert int(pt[0]) == (expected_value if i == index else 0) and not pt[1:].any(), i


def test_expansion_argument_errors(client4096):
    c = client4096
    ct = c.encrypt(np.zeros(N, dtype=np.uint64))
    with pytest.raises(ob.OracleStatus) as e:  # server.cpp:111-114
        c.orc.expand(ct, N + 1, c.elts, c.galois, single=True)
    assert e.value.code == 3
    with pytest.raises(ob.OracleStatus) as e:  # server.cpp:154-158
        c.orc.expand(np.stack([ct, ct]), 100, c.elts, c.galois)
    assert e.value.code == 3


# ------------------------------------------------------------------ ct_reencoder_test.cpp:78 and round trip
def test_reencoder(client4096):
    c = client4096
    assert c.orc.ER == 4  # ct_reencoder_test.cpp:78
    assert ob.Oracle.default(4096, 24).ER == 4 and ob.Oracle.default(8192, 20).ER == 12
    rng = np.random.default_rng(9)
    ct = np.stack([rng.integers(0, c.orc.moduli[j], N, dtype=np.uint64) for j in (0, 1, 0, 1)]).reshape(2, 2, N)
    pts = c.orc.reencode(ct)
    assert pts.shape == (8, N) and int(pts.max()) < (1 << 19)
    assert np.array_equal(c.orc.reencode_decode(pts), ct)


# ------------------------------------------------------------------ PIRServerTest (server_test.cpp:98-260)
def _int_db(params, seed=42):
    rng = np.random.default_rng(seed)
    # test_base.cpp:67-78: 6 random bytes per value ("can't use full size")
    vals = [int(v) for v in rng.integers(0, 1 << 48, params.num_items, dtype=np.int64)]
    return vals


def _server_fixture(dbsize, d=1, plain_bits=20):
    p = oc.create_pir_parameters(dbsize, 7680, d, 4096, plain_bits)
    cl = oc.HarnessClient(p, seed=dbsize + 7)
    vals = _int_db(p)
    db = oc.db_to_ntt(cl.orc, oc.encode_int_db(p, vals))
    return p, cl, vals, db


def test_process_request_single_ct():
    # server_test.cpp:98-121: pt[7] = 1 -> int_db[7] * next_power_two(db_size)
    p, cl, vals, db = _server_fixture(10)
    pt = np.zeros(N, dtype=np.uint64); pt[7] = 1
    reply = cl.orc.process_query(db, p.dimensions, cl.elts, cl.galois, cl.encrypt(pt)[None])
    assert len(reply) == 1
    assert oc.integer_decode(cl.decrypt(reply[0]), p.plain_modulus) == vals[7] * ob.next_power_two(10)


def test_process_request_multi_ct():
    # server_test.cpp:123-151: 5000 items, index 4200 lives in the 2nd ct at slot 104 -> int_db * 1024
    p, cl, vals, db = _server_fixture(5000)
    idx = 4200
    q0 = cl.encrypt(np.zeros(N, dtype=np.uint64))
    pt = np.zeros(N, dtype=np.uint64); pt[idx - N] = 1
    reply = cl.orc.process_query(db, p.dimensions, cl.elts, cl.galois, np.stack([q0, cl.encrypt(pt)]))
    assert len(reply) == 1
    assert oc.integer_decode(cl.decrypt(reply[0]), p.plain_modulus) == vals[idx] * ob.next_power_two(5000 - N)


def test_process_request_zero_query_and_batch():
    # server_test.cpp:153-207
    p, cl, vals, db = _server_fixture(10)
    z = cl.orc.process_query(db, p.dimensions, cl.elts, cl.galois, cl.encrypt(np.zeros(N, dtype=np.uint64))[None])
    assert oc.integer_decode(cl.decrypt(z[0]), p.plain_modulus) == 0
    for idx in (3, 4, 5):
        pt = np.zeros(N, dtype=np.uint64); pt[idx] = 1
        r = cl.orc.process_query(db, p.dimensions, cl.elts, cl.galois, cl.encrypt(pt)[None])
        assert oc.integer_decode(cl.decrypt(r[0]), p.plain_modulus) == vals[idx] * 16


def test_process_request_2dim():
    # server_test.cpp:209-260: 82 items, d=2, dims [10,9], idx 42 -> pt[4], pt[16] = m^-1; reply = 2*ER cts
    p, cl, vals, db = _server_fixture(82, 2)
    assert p.dimensions == [10, 9] and oc.calculate_indices(p, 42) == [4, 6]
    m_inv = pow(ob.next_power_two(19), -1, p.plain_modulus)
    pt = np.zeros(N, dtype=np.uint64); pt[4] = m_inv; pt[16] = m_inv
    reply = cl.orc.process_query(db, p.dimensions, cl.elts, cl.galois, cl.encrypt(pt)[None])
    assert len(reply) == cl.orc.ER * 2
    res, budget = cl.process_reply(reply, with_budget=True)
    assert oc.integer_decode(res, p.plain_modulus) == vals[42]
    assert budget > 0


# ------------------------------------------------------------------ database_test.cpp
def _selection_vector(cl, dims, indices):
    cts = []
    for d, dim in enumerate(dims):
        for i in range(dim):
            pt = np.zeros(cl.orc.N, dtype=np.uint64)
            if i == indices[d]:
                pt[0] = 1
            cts.append(cl.encrypt(pt))
    return np.stack(cts)


def test_db_multiply_integer_dot_product():
    # database_test.cpp:155-178: sv encrypts -n/2.. ; result = sum v[i]*db[i]
    p = oc.create_pir_parameters(10, 0, 1, 4096, 20)
    cl = oc.HarnessClient(p, seed=3)
    rng = np.random.default_rng(8)
    vals = [int(v) for v in rng.integers(0, 1 << 20, 10)]
    db = oc.db_to_ntt(cl.orc, oc.encode_int_db(p, vals))
    v = list(range(-5, 5))
    sv = np.stack([cl.encrypt(oc.integer_encode(x, N, p.plain_modulus)) for x in v])
    out, sv_after = cl.orc.db_multiply(db, p.dimensions, sv)
    assert len(out) == 1
    assert oc.integer_decode(cl.decrypt(out[0]), p.plain_modulus) == sum(a * b for a, b in zip(v, vals))
    # database.cpp:190 — selection vector is transformed to NTT form in place
    assert np.array_equal(sv_after[0], cl.orc.ct_to_ntt(sv[0]))


def test_db_multiply_wrong_selection_vector_size():
    # database_test.cpp:180-219
    p = oc.create_pir_parameters(100, 0, 2, 4096, 20)
    o = ob.Oracle.default(4096, 20)
    db = np.zeros((100, o.k, N), dtype=np.uint64)
    for n_sv in (19, 21):
        with pytest.raises(ob.OracleStatus) as e:
            o.db_multiply(db, p.dimensions, np.zeros((n_sv, 2, o.k, N), dtype=np.uint64))
        assert e.value.code == 3


@pytest.mark.parametrize("n,bits,dbsize,d,idx", [(4096, 16, 10, 1, 7), (4096, 16, 16, 2, 11), (4096, 16, 16, 2, 0),
                                                  (4096, 16, 16, 2, 15), (4096, 16, 82, 2, 42), (8192, 20, 27, 3, 2),
                                                  (8192, 20, 117, 3, 17)])
def test_multiply_multi_dim_strings(n, bits, dbsize, d, idx):
    # database_test.cpp:343-388 (CTDecomp arm)
    p = oc.create_pir_parameters(dbsize, 0, d, n, bits)
    cl = oc.HarnessClient(p, seed=11)
    rng = np.random.default_rng(42)
    items = [rng.integers(0, 256, p.bytes_per_item, dtype=np.uint8).tobytes() for _ in range(dbsize)]
    db = oc.db_to_ntt(cl.orc, oc.encode_string_db(p, items))
    sv = _selection_vector(cl, p.dimensions, oc.calculate_indices(p, idx))
    out, _ = cl.orc.db_multiply(db, p.dimensions, sv)
    assert len(out) == (2 * cl.orc.ER) ** (d - 1)
    res = cl.process_reply(out)
    assert ob.string_decode(res, cl.orc.ptb, p.bytes_per_item) == items[idx]


# ------------------------------------------------------------------ correctness_test.cpp:107-113 (decomposition arm)
@pytest.mark.parametrize("n,bits,elem,bpc,dbsize,d,indices",
                         [(4096, 24, 0, 0, 10, 1, [0]), (4096, 24, 0, 10, 9, 2, [1, 5]),
                          (4096, 24, 0, 6, 500, 2, [9, 125]), (4096, 24, 64, 10, 1200, 1, [0, 80, 81, 123, 777, 1199]),
                          (4096, 24, 289, 10, 1200, 1, [0, 47, 777, 1199])])
def test_end_to_end_correctness(n, bits, elem, bpc, dbsize, d, indices):
    p = oc.create_pir_parameters(dbsize, elem, d, n, bits, bpc)
    cl = oc.HarnessClient(p, seed=5)
    rng = np.random.default_rng(42)
    items = [rng.integers(0, 256, p.bytes_per_item, dtype=np.uint8).tobytes() for _ in range(dbsize)]
    db = oc.db_to_ntt(cl.orc, oc.encode_string_db(p, items))
    replies = [cl.orc.process_query(db, p.dimensions, cl.elts, cl.galois, cl.create_query(i)) for i in indices]
    got = cl.process_response_strings(indices, replies)
    assert got == [items[i] for i in indices]


# ------------------------------------------------------------------ ciphertext-multiplication mode (database.cpp:202-211)
def test_behz_auxiliary_bases():
    # [SEAL RNSTool::initialize]: m_sk, gamma, then |q| primes for B from get_primes(N, 61, .): descending from 2^61,
    # all = 1 mod 2N; m_tilde = 2^32.  4096: |q| = 2 data primes, 8192: 4.
    for n, k in ((4096, 2), (8192, 4)):
        o = ob.Oracle.default(n, 20)
        m_sk, B = o.rns_bases()
        assert len(B) == k
        aux = [m_sk] + B
        assert all(ob.is_prime(p) and p % (2 * n) == 1 and p.bit_length() == 61 for p in aux)
        assert aux[0] > aux[1] and all(B[i] > B[i + 1] for i in range(k - 1))
        assert m_sk == max(p for p in range((1 << 61) - 2 * n + 1, (1 << 61) - 200 * n, -2 * n) if ob.is_prime(p))


def test_bfv_multiply_and_relinearize():
    """Evaluator::multiply + relinearize_inplace at the decrypt level: the product of two plaintext polynomials
    mod (x^N + 1, t), through a size-3 ciphertext, through relinearization, and through a size-4 one."""
    p = oc.create_pir_parameters(10, 0, 1, 8192, 20, use_ciphertext_multiplication=True)
    cl = oc.HarnessClient(p, seed=21)
    n, t = 8192, p.plain_modulus
    rng = np.random.default_rng(5)
    a = np.zeros(n, dtype=np.uint64); b = np.zeros(n, dtype=np.uint64)
    ia, ib = rng.choice(n, 40, replace=False), rng.choice(n, 3, replace=False)
    a[ia] = rng.integers(1, t, 40); b[ib] = rng.integers(1, t, 3)
    want = np.zeros(n, dtype=object)
    for i in ia:
        for j in ib:
            d, v = int(i + j), int(a[i]) * int(b[j])
            if d >= n:
                d, v = d - n, -v
            want[d] = (want[d] + v) % t
    ca, cb = cl.encrypt(a), cl.encrypt(b)
    prod = cl.orc.bfv_multiply(ca, cb)
    assert prod.shape == (3, cl.orc.k, n)
    got, budget3 = cl.orc.decrypt_polys(cl.keys, prod, True)
    assert [int(x) for x in got] == [int(x) for x in want] and budget3 > 60
    assert np.array_equal(cl.orc.bfv_multiply(cb, ca), prod)  # symmetric in its operands
    rel = cl.orc.relinearize(prod, cl.relin)
    got, budget2 = cl.decrypt(rel, True)
    assert [int(x) for x in got] == [int(x) for x in want] and budget2 >= budget3 - 2
    one = np.zeros(n, dtype=np.uint64); one[0] = 1
    prod4 = cl.orc.bfv_multiply(prod, cl.encrypt(one))
    assert prod4.shape == (4, cl.orc.k, n)
    assert [int(x) for x in cl.orc.decrypt_polys(cl.keys, prod4)] == [int(x) for x in want]


@pytest.mark.parametrize("n,bits,dbsize,d,idx", [(4096, 16, 10, 1, 7), (4096, 16, 16, 2, 11), (4096, 16, 16, 2, 0),
                                                  (4096, 16, 16, 2, 15), (4096, 16, 82, 2, 42), (8192, 20, 27, 3, 2),
                                                  (8192, 20, 117, 3, 17)])
def test_multiply_multi_dim_strings_ct_multiply(n, bits, dbsize, d, idx):
    # database_test.cpp:343-388, CTMultiply arm: relinearization keys given, ONE reply ciphertext of size 2
    p = oc.create_pir_parameters(dbsize, 0, d, n, bits, use_ciphertext_multiplication=True)
    cl = oc.HarnessClient(p, seed=11)
    rng = np.random.default_rng(42)
    items = [rng.integers(0, 256, p.bytes_per_item, dtype=np.uint8).tobytes() for _ in range(dbsize)]
    db = oc.db_to_ntt(cl.orc, oc.encode_string_db(p, items))
    sv = _selection_vector(cl, p.dimensions, oc.calculate_indices(p, idx))
    out = cl.orc.db_multiply_ct(db, p.dimensions, sv, cl.relin)
    assert out.shape == (2, cl.orc.k, n)
    res, budget = cl.process_reply_ct(out, with_budget=True)
    assert budget > 0
    assert ob.string_decode(res, cl.orc.ptb, p.bytes_per_item) == items[idx]
    if d == 2:  # server.cpp:185-190 without relinearization keys: the reply keeps its third polynomial
        out3 = cl.orc.db_multiply_ct(db, p.dimensions, sv, None)
        assert out3.shape == (3, cl.orc.k, n)
        assert ob.string_decode(cl.process_reply_ct(out3), cl.orc.ptb, p.bytes_per_item) == items[idx]
        with pytest.raises(ob.OracleStatus) as e:  # database.cpp:297-300
            cl.orc.db_multiply_ct(db, p.dimensions, sv[:-1], cl.relin)
        assert e.value.code == 3


def test_process_request_2dim_ct_multiply():
    # server_test.cpp:209-260 with GetParam() == true: reply is one ciphertext of size 2 ("Were relin keys used?")
    p = oc.create_pir_parameters(82, 7680, 2, 4096, 20, use_ciphertext_multiplication=True)
    cl = oc.HarnessClient(p, seed=89)
    vals = _int_db(p)
    db = oc.db_to_ntt(cl.orc, oc.encode_int_db(p, vals))
    m_inv = pow(ob.next_power_two(19), -1, p.plain_modulus)
    pt = np.zeros(N, dtype=np.uint64); pt[4] = m_inv; pt[16] = m_inv
    reply = cl.orc.process_query_ct(db, p.dimensions, cl.elts, cl.galois, cl.encrypt(pt)[None], cl.relin)
    assert reply.shape == (2, cl.orc.k, N)
    assert oc.integer_decode(cl.process_reply_ct(reply), p.plain_modulus) == vals[42]


@pytest.mark.parametrize("n,bits,elem,bpc,dbsize,d,indices",
                         [(4096, 24, 0, 0, 10, 1, [0]), (4096, 16, 0, 10, 9, 2, [1, 5]),
                          (4096, 16, 0, 6, 500, 2, [9, 125]), (8192, 42, 0, 0, 87, 2, [5, 33, 86]),
                          (4096, 16, 64, 10, 1200, 1, [0, 80, 81, 123, 777, 1199]),
                          (4096, 16, 289, 10, 1200, 1, [0, 47, 777, 1199])])
def test_end_to_end_correctness_ct_multiply(n, bits, elem, bpc, dbsize, d, indices):
    # correctness_test.cpp:94-105: the six use_ciphertext_multiplication == true cases
    p = oc.create_pir_parameters(dbsize, elem, d, n, bits, bpc, use_ciphertext_multiplication=True)
    cl = oc.HarnessClient(p, seed=5)
    rng = np.random.default_rng(42)
    items = [rng.integers(0, 256, p.bytes_per_item, dtype=np.uint8).tobytes() for _ in range(dbsize)]
    db = oc.db_to_ntt(cl.orc, oc.encode_string_db(p, items))
    replies = [cl.orc.process_query_ct(db, p.dimensions, cl.elts, cl.galois, cl.create_query(i), cl.relin)
               for i in indices]
    assert all(r.shape[0] == 2 for r in replies)
    assert cl.process_response_strings(indices, replies) == [items[i] for i in indices]


# ------------------------------------------------------------------ the oracle frozen against itself
def test_oracle_outputs_match_the_committed_digests():
    """tests/golden/oracle_digests.json: SHA-256 of the oracle's substitution / shift / expansion / reply for inputs
    derived from SHAKE-256 (pure integer pipeline).  The oracle defines "bit-exact" for the CUDA path, so a change in
    oracle/ that moves a single limb must be deliberate: regenerate with tests/golden/make_oracle_digests.py."""
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_oracle_digests", os.path.join(here, "golden",
                                                                                      "make_oracle_digests.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    with open(os.path.join(here, "golden", "oracle_digests.json")) as f:
        want = json.load(f)
    for name, *args in mod.CASES:
        assert mod.run_case(name, *args) == want[name], name
