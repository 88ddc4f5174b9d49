"""Wire format: the hand-rolled protobuf codec against the python protobuf runtime, BLAKE2b/BLAKE2xb
against hashlib and an independent implementation, SEAL-object save/load round trips and seed expansion.

The SEAL 3.5.6 object layout itself cannot be byte-verified here (no SEAL in this image); see pir_b200/cpp/wire.hpp."""
import ctypes as C
import hashlib
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pir_b200", "lib", "libpirb_wire.so")

u8p, u32p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        subprocess.check_call(["make", "-C", ROOT, "wire"])
    L = C.CDLL(LIB)
    L.pirw_last_error.restype = C.c_char_p
    return L


def take(lib, out, n):
    b = C.string_at(out, n.value)
    lib.pirw_free(out)
    return b


def buf(b):
    return (C.c_uint8 * max(1, len(b))).from_buffer_copy(b if len(b) else b"\0")


# ---------------------------------------------------------------------------------------------------------------
# protobuf framing vs the python protobuf runtime (messages restated from pir/proto/payload.proto:20-69)
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def pb():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name="payload.proto", package="pir", syntax="proto3")

    def msg(name, fields):
        m = fd.message_type.add(name=name)
        for fname, num, typ, label, tname in fields:
            f = m.field.add(name=fname, number=num, type=typ, label=label)
            if tname:
                f.type_name = tname

    REP, OPT = F.LABEL_REPEATED, F.LABEL_OPTIONAL
    msg("Ciphertexts", [("ct", 1, F.TYPE_BYTES, REP, None)])
    msg("Request", [("query", 1, F.TYPE_MESSAGE, REP, ".pir.Ciphertexts"), ("galois_keys", 2, F.TYPE_BYTES, OPT, None),
                    ("relin_keys", 3, F.TYPE_BYTES, OPT, None)])
    msg("Response", [("reply", 1, F.TYPE_MESSAGE, REP, ".pir.Ciphertexts")])
    msg("PIRParameters", [("num_items", 1, F.TYPE_UINT64, OPT, None), ("num_pt", 4, F.TYPE_UINT64, OPT, None),
                          ("dimensions", 2, F.TYPE_UINT32, REP, None),
                          ("encryption_parameters", 3, F.TYPE_BYTES, OPT, None),
                          ("bytes_per_item", 5, F.TYPE_UINT32, OPT, None),
                          ("items_per_plaintext", 6, F.TYPE_UINT32, OPT, None),
                          ("bits_per_coeff", 7, F.TYPE_UINT32, OPT, None),
                          ("use_ciphertext_multiplication", 8, F.TYPE_BOOL, OPT, None)])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("pir." + n))  # noqa: E731
    return {n: get(n) for n in ("Ciphertexts", "Request", "Response", "PIRParameters")}


def roundtrip(lib, kind, data):
    out, n = u8p(), C.c_size_t()
    rc = lib.pirw_proto_roundtrip(kind, buf(data), C.c_size_t(len(data)), C.byref(out), C.byref(n))
    if rc:
        return rc, None
    return 0, take(lib, out, n)


def test_request_matches_protobuf_runtime(lib, pb):
    rng = np.random.default_rng(1)
    for n_q, n_ct, ct_len, gk_len, rk_len in [(1, 1, 300, 5000, 70), (3, 2, 129, 0, 0), (0, 0, 0, 10, 0),
                                              (2, 1, 0, 3, 200000), (1, 3, 70000, 1, 1)]:
        cts = rng.integers(0, 256, size=n_q * n_ct * ct_len, dtype=np.uint8).tobytes()
        gk = rng.integers(0, 256, size=gk_len, dtype=np.uint8).tobytes()
        rk = rng.integers(0, 256, size=rk_len, dtype=np.uint8).tobytes()
        ref = pb["Request"]()
        for q in range(n_q):
            c = ref.query.add()
            for i in range(n_ct):
                c.ct.append(cts[(q * n_ct + i) * ct_len:(q * n_ct + i + 1) * ct_len])
        ref.galois_keys, ref.relin_keys = gk, rk
        want = ref.SerializeToString(deterministic=True)
        out, n = u8p(), C.c_size_t()
        assert lib.pirw_request_build(buf(cts), n_q, n_ct, C.c_size_t(ct_len), buf(gk), C.c_size_t(len(gk)), buf(rk),
                                      C.c_size_t(len(rk)), C.byref(out), C.byref(n)) == 0
        got = take(lib, out, n)
        assert got == want
        # and the parser: protobuf-runtime bytes -> parse -> serialize is the identity
        rc, again = roundtrip(lib, 1, want)
        assert rc == 0 and again == want
        back = pb["Request"]()
        back.ParseFromString(again)
        assert back == ref


def test_response_and_ciphertexts_match_protobuf_runtime(lib, pb):
    rng = np.random.default_rng(2)
    resp = pb["Response"]()
    for q in range(3):
        c = resp.reply.add()
        for i in range(q + 1):
            c.ct.append(rng.integers(0, 256, size=1000 + 77 * i, dtype=np.uint8).tobytes())
    resp.reply.add()  # an empty reply is still framed
    want = resp.SerializeToString(deterministic=True)
    rc, got = roundtrip(lib, 2, want)
    assert rc == 0 and got == want
    one = resp.reply[2].SerializeToString(deterministic=True)
    rc, got = roundtrip(lib, 0, one)
    assert rc == 0 and got == one


def test_pir_parameters_match_protobuf_runtime(lib, pb):
    cases = [dict(num_items=65536, num_pt=1639, dimensions=[41, 40], encryption_parameters=b"\x5e\xa1" + bytes(100),
                  bytes_per_item=288, items_per_plaintext=40, bits_per_coeff=0, use_ciphertext_multiplication=False),
             dict(num_items=1 << 40, num_pt=(1 << 33) + 5, dimensions=[1 << 31, 3, 0, 300], encryption_parameters=b"",
                  bytes_per_item=0, items_per_plaintext=1, bits_per_coeff=19, use_ciphertext_multiplication=True),
             dict(num_items=0, num_pt=0, dimensions=[], encryption_parameters=b"", bytes_per_item=0,
                  items_per_plaintext=0, bits_per_coeff=0, use_ciphertext_multiplication=False)]
    for c in cases:
        ref = pb["PIRParameters"](**c)
        want = ref.SerializeToString(deterministic=True)
        dims = np.array(c["dimensions"], dtype=np.uint32)
        out, n = u8p(), C.c_size_t()
        ep = c["encryption_parameters"]
        assert lib.pirw_params_build(C.c_uint64(c["num_items"]), C.c_uint64(c["num_pt"]),
                                     dims.ctypes.data_as(u32p), len(dims), buf(ep), C.c_size_t(len(ep)),
                                     c["bytes_per_item"], c["items_per_plaintext"], c["bits_per_coeff"],
                                     int(c["use_ciphertext_multiplication"]), C.byref(out), C.byref(n)) == 0
        assert take(lib, out, n) == want
        rc, got = roundtrip(lib, 3, want)
        assert rc == 0 and got == want


def test_parser_skips_unknown_fields_and_rejects_garbage(lib, pb):
    ref = pb["Request"]()
    ref.query.add().ct.append(b"abc")
    ref.galois_keys = b"k"
    good = ref.SerializeToString()
    # unknown varint field 15, unknown fixed64 field 14, unknown length-delimited field 13, unknown fixed32 field 12
    extra = bytes([15 << 3 | 0, 0x96, 0x01]) + bytes([14 << 3 | 1]) + bytes(8) + bytes([13 << 3 | 2, 2, 7, 7]) + \
        bytes([12 << 3 | 5]) + bytes(4)
    rc, got = roundtrip(lib, 1, extra + good)
    assert rc == 0 and got == good
    for bad in [good[:-1], b"\x0a\xff\xff\xff\xff\x0f", b"\x0b", b"\x00\x00", b"\x0a\x05abc"]:
        rc, _ = roundtrip(lib, 1, bad)
        assert rc == 3, bad


# ---------------------------------------------------------------------------------------------------------------
# BLAKE2b / BLAKE2xb
# ---------------------------------------------------------------------------------------------------------------
_IV = [0x6a09e667f3bcc908, 0xbb67ae8584caa73b, 0x3c6ef372fe94f82b, 0xa54ff53a5f1d36f1,
       0x510e527fade682d1, 0x9b05688c2b3e6c1f, 0x1f83d9abfb41bd6b, 0x5be0cd19137e2179]
_M64 = (1 << 64) - 1


def _sigma():
    # RFC 7693 section 2.7, rounds 10 and 11 repeat rounds 0 and 1
    s = ["0123456789abcdef", "ea489fd61c02b753", "b8c052fdae367194", "7931dcbe265a40f8", "905724afe1bc683d",
         "2c6a0b834d75fe19", "c51fed4a0763928b", "db7ec13950f4862a", "6fe9b308c2d714a5", "a2847615fb9e3cd0"]
    rows = [[int(ch, 16) for ch in r] for r in s]
    return rows + rows[:2]


def _py_blake2b_param(param_block, data, key=b"", outlen=64):
    """Straight-from-the-RFC BLAKE2b with an explicit 64-byte parameter block (independent of wire.hpp)."""
    sig = _sigma()
    h = [_IV[i] ^ struct.unpack_from("<Q", param_block, 8 * i)[0] for i in range(8)]
    if key:
        data = key.ljust(128, b"\0") + data
    blocks = [data[i:i + 128] for i in range(0, len(data), 128)] or [b""]

    def rotr(x, n):
        return ((x >> n) | (x << (64 - n))) & _M64

    t = 0
    for bi, blk in enumerate(blocks):
        last = bi == len(blocks) - 1
        t += len(blk)
        m = list(struct.unpack("<16Q", blk.ljust(128, b"\0")))
        v = h[:] + _IV[:]
        v[12] ^= t & _M64
        v[13] ^= t >> 64
        if last:
            v[14] ^= _M64
        for r in range(12):
            s = sig[r]
            for i, (a, b, c, d) in enumerate([(0, 4, 8, 12), (1, 5, 9, 13), (2, 6, 10, 14), (3, 7, 11, 15),
                                              (0, 5, 10, 15), (1, 6, 11, 12), (2, 7, 8, 13), (3, 4, 9, 14)]):
                v[a] = (v[a] + v[b] + m[s[2 * i]]) & _M64
                v[d] = rotr(v[d] ^ v[a], 32)
                v[c] = (v[c] + v[d]) & _M64
                v[b] = rotr(v[b] ^ v[c], 24)
                v[a] = (v[a] + v[b] + m[s[2 * i + 1]]) & _M64
                v[d] = rotr(v[d] ^ v[a], 16)
                v[c] = (v[c] + v[d]) & _M64
                v[b] = rotr(v[b] ^ v[c], 63)
        h = [h[i] ^ v[i] ^ v[i + 8] for i in range(8)]
    return struct.pack("<8Q", *h)[:outlen]


def _py_blake2xb(outlen, data, key):
    root_p = struct.pack("<BBBBIIIBB", 64, len(key), 1, 1, 0, 0, outlen, 0, 0).ljust(64, b"\0")
    root = _py_blake2b_param(root_p, data, key, 64)
    out = b""
    i = 0
    while len(out) < outlen:
        n = min(64, outlen - len(out))
        p = struct.pack("<BBBBIIIBB", n, 0, 0, 0, 64, i, outlen, 0, 64).ljust(64, b"\0")
        out += _py_blake2b_param(p, root, b"", n)
        i += 1
    return out


def c_blake(lib, fn, outlen, data, key):
    out = (C.c_uint8 * outlen)()
    getattr(lib, fn)(out, C.c_size_t(outlen), buf(data), C.c_size_t(len(data)), buf(key), C.c_size_t(len(key)))
    return bytes(out)


def test_blake2b_matches_hashlib(lib):
    rng = np.random.default_rng(3)
    # RFC 7693 appendix A: BLAKE2b-512("abc")
    assert c_blake(lib, "pirw_blake2b", 64, b"abc", b"").hex().startswith("ba80a53f981c4d0d6a2797b69f12f6e9")
    for n in [0, 1, 3, 64, 127, 128, 129, 255, 256, 257, 1000, 4096]:
        data = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
        for outlen in (32, 64, 20):
            for klen in (0, 16, 64):
                key = rng.integers(0, 256, size=klen, dtype=np.uint8).tobytes()
                want = hashlib.blake2b(data, digest_size=outlen, key=key).digest()
                assert c_blake(lib, "pirw_blake2b", outlen, data, key) == want
                p = struct.pack("<BBBBIIIBB", outlen, klen, 1, 1, 0, 0, 0, 0, 0).ljust(64, b"\0")
                assert _py_blake2b_param(p, data, key, outlen) == want  # pins the independent implementation too


def test_blake2xb_matches_independent_implementation(lib):
    rng = np.random.default_rng(4)
    for outlen in (1, 63, 64, 65, 200, 4096):
        for n in (0, 8, 130):
            data = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
            key = rng.integers(0, 256, size=64, dtype=np.uint8).tobytes()
            assert c_blake(lib, "pirw_blake2xb", outlen, data, key) == _py_blake2xb(outlen, data, key)
    # the expansion blocks are ordinary BLAKE2b calls with tree parameters hashlib can express except depth = 0;
    # cross-check the root hash (node_offset's upper half is the XOF length) with hashlib
    key, data, outlen = bytes(range(64)), struct.pack("<Q", 0), 4096
    root = hashlib.blake2b(data, digest_size=64, key=key, fanout=1, depth=1, node_offset=outlen << 32).digest()
    p = struct.pack("<BBBBIIIBB", 64, 64, 1, 1, 0, 0, outlen, 0, 0).ljust(64, b"\0")
    assert _py_blake2b_param(p, data, key, 64) == root


# ---------------------------------------------------------------------------------------------------------------
# SEAL objects: layout constants, round trips, seed expansion
# ---------------------------------------------------------------------------------------------------------------
MODS = np.array([0xffffee001, 0xffffc4001, 0x1ffffe0001], dtype=np.uint64)  # BFVDefault(4096)
N, T = 4096, 0xFC001


def rand_limbs(rng, mods, prefix):
    return np.ascontiguousarray(np.stack([rng.integers(0, int(q), size=prefix + (N,), dtype=np.uint64) for q in mods],
                                         axis=len(prefix)))


def test_ciphertext_save_load_roundtrip_and_layout(lib):
    rng = np.random.default_rng(5)
    k = 2
    limbs = rand_limbs(rng, MODS[:k], (2,))
    pid = np.zeros(4, dtype=np.uint64)
    lib.pirw_parms_id(N, MODS.ctypes.data_as(u64p), k, C.c_uint64(T), pid.ctypes.data_as(u64p))
    words = np.array([1, N] + [int(q) for q in MODS[:k]] + [T], dtype=np.uint64)
    assert pid.tobytes() == hashlib.blake2b(words.tobytes(), digest_size=32).digest()
    out, n = u8p(), C.c_size_t()
    assert lib.pirw_ct_save(limbs.ctypes.data_as(u64p), 2, N, k, pid.ctypes.data_as(u64p), 0, None, C.byref(out),
                            C.byref(n)) == 0
    blob = take(lib, out, n)
    # SEAL 3.5 header {magic 0xA15E, header size 0x10, version 3.5, compr none, reserved, u64 total size}; members;
    # nested IntArray header + count + limbs
    assert len(blob) == 16 + 32 + 1 + 8 * 3 + 8 + 16 + 8 + limbs.nbytes
    magic, hsize, vmaj, vmin, compr, reserved, size = struct.unpack_from("<HBBBBHQ", blob, 0)
    assert (magic, hsize, vmaj, vmin, compr, reserved, size) == (0xA15E, 0x10, 3, 5, 0, 0, len(blob))
    assert blob[16:48] == pid.tobytes() and blob[48] == 0
    assert struct.unpack_from("<QQQd", blob, 49) == (2, N, k, 1.0)
    assert struct.unpack_from("<HBBBBHQQ", blob, 81) == (0xA15E, 0x10, 3, 5, 0, 0, 16 + 8 + limbs.nbytes, limbs.size)
    assert blob[105:] == limbs.tobytes()
    got = np.zeros_like(limbs)
    pid2 = np.zeros(4, dtype=np.uint64)
    ntt, seeded = C.c_int(), C.c_int()
    assert lib.pirw_ct_load(buf(blob), C.c_size_t(len(blob)), N, MODS.ctypes.data_as(u64p), k,
                            got.ctypes.data_as(u64p), pid2.ctypes.data_as(u64p), C.byref(ntt), C.byref(seeded)) == 0
    assert np.array_equal(got, limbs) and np.array_equal(pid, pid2) and ntt.value == 0 and seeded.value == 0
    # the loader also accepts the legacy SEAL 3.4 header {magic, 0, compr, u32 size, u64 reserved}, as SEAL 3.5's
    # LoadHeader does, and rejects a foreign version or header size
    def legacy(b, off):
        size = struct.unpack_from("<Q", b, off + 8)[0]
        return b[:off] + struct.pack("<HBBIQ", 0xA15E, 0, 0, size, 0) + b[off + 16:]
    old = legacy(legacy(blob, 81), 0)
    got2 = np.zeros_like(limbs)
    assert lib.pirw_ct_load(buf(old), C.c_size_t(len(old)), N, MODS.ctypes.data_as(u64p), k,
                            got2.ctypes.data_as(u64p), pid2.ctypes.data_as(u64p), C.byref(ntt), C.byref(seeded)) == 0
    assert np.array_equal(got2, limbs)
    for off, val in ((2, 0x18), (3, 4), (4, 6)):
        bad = bytearray(blob)
        bad[off] = val
        assert lib.pirw_ct_load(buf(bytes(bad)), C.c_size_t(len(bad)), N, MODS.ctypes.data_as(u64p), k,
                                got2.ctypes.data_as(u64p), pid2.ctypes.data_as(u64p), C.byref(ntt), C.byref(seeded)) != 0
    # malformed inputs are InvalidArgument (serialization.h:113-115), never a crash
    bad_magic = b"\x00" + blob[1:]
    too_big = bytearray(blob)
    struct.pack_into("<Q", too_big, 105, int(MODS[0]))  # first limb == q_0
    wrong_n = bytearray(blob)
    struct.pack_into("<Q", wrong_n, 57, 2048)
    for bad in [blob[:50], blob[:-8], bad_magic, bytes(too_big), bytes(wrong_n), b""]:
        assert lib.pirw_ct_load(buf(bad), C.c_size_t(len(bad)), N, MODS.ctypes.data_as(u64p), k,
                                got.ctypes.data_as(u64p), pid2.ctypes.data_as(u64p), C.byref(ntt),
                                C.byref(seeded)) == 3
        assert lib.pirw_last_error()


def py_sample_poly_uniform(seed_words, mods):
    """util::sample_poly_uniform over BlakePRNG restated in python on top of the independent BLAKE2xb."""
    key = np.array(seed_words, dtype=np.uint64).tobytes()
    state = {"buf": b"", "pos": 0, "ctr": 0}

    def gen():
        if state["pos"] == len(state["buf"]):
            state["buf"] = _py_blake2xb(4096, struct.pack("<Q", state["ctr"]), key)
            state["ctr"] += 1
            state["pos"] = 0
        v = struct.unpack_from("<I", state["buf"], state["pos"])[0]
        state["pos"] += 4
        return v

    out = []
    maxr = 0x7FFFFFFFFFFFFFFF
    for q in mods:
        lim = maxr - (maxr % int(q)) - 1
        row = []
        while len(row) < 64:  # only the first coefficients (pure python BLAKE2 is slow)
            a, b = gen(), gen()
            r = (a << 31) | (b >> 1)
            if r < lim:
                row.append(r % int(q))
        out.append(row)
        # the C sampler continues in the same stream for the next modulus, so only modulus 0 is comparable here
        break
    return out


def test_seed_expansion_matches_python_restatement(lib):
    seed = np.arange(1, 9, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    dst = np.zeros((3, N), dtype=np.uint64)
    lib.pirw_sample_poly_uniform(seed.ctypes.data_as(u64p), N, MODS.ctypes.data_as(u64p), 3, dst.ctypes.data_as(u64p))
    for j in range(3):
        assert dst[j].max() < MODS[j]
    want = py_sample_poly_uniform(seed, MODS)
    assert [int(x) for x in dst[0, :64]] == want[0]
    # uniformity smoke check: mean within 2% of q/2
    assert abs(float(dst[0].mean()) / float(MODS[0]) - 0.5) < 0.02


def test_galois_keys_roundtrip_plain_and_seeded(lib):
    rng = np.random.default_rng(6)
    k = 2
    elts = np.array([(N >> i) + 1 for i in range(12)], dtype=np.uint32)
    limbs = rand_limbs(rng, MODS, (len(elts), k, 2))
    out, n = u8p(), C.c_size_t()
    assert lib.pirw_galois_keys_save(N, MODS.ctypes.data_as(u64p), 3, C.c_uint64(T), elts.ctypes.data_as(u32p),
                                     len(elts), limbs.ctypes.data_as(u64p), None, C.byref(out), C.byref(n)) == 0
    blob = take(lib, out, n)
    per_key = 16 + 32 + 1 + 24 + 8 + 16 + 8 + 2 * 3 * N * 8
    assert len(blob) == 16 + 32 + 8 + N * 8 + len(elts) * k * per_key
    got_e = np.zeros(64, dtype=np.uint32)
    got_l = np.zeros_like(limbs)
    cnt = C.c_uint32()
    assert lib.pirw_galois_keys_load(buf(blob), C.c_size_t(len(blob)), N, MODS.ctypes.data_as(u64p), 3, C.c_uint64(T),
                                     len(elts), got_e.ctypes.data_as(u32p), got_l.ctypes.data_as(u64p),
                                     C.byref(cnt)) == 0
    order = np.argsort(elts)  # slots come back in index order
    assert cnt.value == len(elts) and np.array_equal(got_e[:cnt.value], elts[order])
    assert np.array_equal(got_l, limbs[order])
    # seed-compressed keys (what client.cpp:47-54 sends): ~half the bytes; the second polynomial is the seed expansion
    seeds = rng.integers(0, 1 << 63, size=(len(elts), k, 8), dtype=np.uint64)
    assert lib.pirw_galois_keys_save(N, MODS.ctypes.data_as(u64p), 3, C.c_uint64(T), elts.ctypes.data_as(u32p),
                                     len(elts), limbs.ctypes.data_as(u64p), seeds.ctypes.data_as(u64p), C.byref(out),
                                     C.byref(n)) == 0
    sblob = take(lib, out, n)
    assert len(sblob) == len(blob) - len(elts) * k * (3 * N * 8 - 64)
    got_s = np.zeros_like(limbs)
    assert lib.pirw_galois_keys_load(buf(sblob), C.c_size_t(len(sblob)), N, MODS.ctypes.data_as(u64p), 3,
                                     C.c_uint64(T), len(elts), got_e.ctypes.data_as(u32p), got_s.ctypes.data_as(u64p),
                                     C.byref(cnt)) == 0
    assert np.array_equal(got_s[:, :, 0], limbs[order][:, :, 0])
    exp = np.zeros((3, N), dtype=np.uint64)
    for e in (0, 5, 11):
        for j in range(k):
            lib.pirw_sample_poly_uniform(seeds[order][e, j].ctypes.data_as(u64p), N, MODS.ctypes.data_as(u64p), 3,
                                         exp.ctypes.data_as(u64p))
            assert np.array_equal(got_s[e, j, 1], exp)
    # truncated / corrupted key blobs are rejected
    for bad in [blob[:1000], sblob[:-1], blob[:16] + bytes(40)]:
        assert lib.pirw_galois_keys_load(buf(bad), C.c_size_t(len(bad)), N, MODS.ctypes.data_as(u64p), 3,
                                         C.c_uint64(T), len(elts), got_e.ctypes.data_as(u32p),
                                         got_s.ctypes.data_as(u64p), C.byref(cnt)) == 3


def test_encryption_parameters_roundtrip(lib):
    out, n = u8p(), C.c_size_t()
    assert lib.pirw_encryption_parameters_save(N, MODS.ctypes.data_as(u64p), 3, C.c_uint64(T), C.byref(out),
                                               C.byref(n)) == 0
    blob = take(lib, out, n)
    assert len(blob) == 16 + 1 + 8 + 8 + 4 * 24
    gn, gm, gt = C.c_uint32(), C.c_uint32(), C.c_uint64()
    mods = np.zeros(8, dtype=np.uint64)
    assert lib.pirw_encryption_parameters_load(buf(blob), C.c_size_t(len(blob)), C.byref(gn), mods.ctypes.data_as(u64p),
                                               8, C.byref(gm), C.byref(gt)) == 0
    assert (gn.value, gm.value, gt.value) == (N, 3, T) and np.array_equal(mods[:3], MODS)
    assert lib.pirw_encryption_parameters_load(buf(blob[:30]), C.c_size_t(30), C.byref(gn), mods.ctypes.data_as(u64p),
                                               8, C.byref(gm), C.byref(gt)) == 3


# ---------------------------------------------------------------------------------------------------------------
# untrusted input: a server parses these bytes straight off the network
# ---------------------------------------------------------------------------------------------------------------
def test_mutated_objects_never_crash_the_loaders(lib):
    """Random byte flips, truncations and splices of valid objects must come back as OK or InvalidArgument (3) —
    never a crash, an over-read or an unbounded allocation."""
    rng = np.random.default_rng(11)
    k = 2
    pid = np.zeros(4, dtype=np.uint64)
    out, n = u8p(), C.c_size_t()
    limbs = rand_limbs(rng, MODS[:k], (2,))
    assert lib.pirw_ct_save(limbs.ctypes.data_as(u64p), 2, N, k, pid.ctypes.data_as(u64p), 0, None, C.byref(out),
                            C.byref(n)) == 0
    ct_blob = take(lib, out, n)
    elts = np.array([N + 1, N // 2 + 1, 3], dtype=np.uint32)
    klimbs = rand_limbs(rng, MODS, (len(elts), k, 2))
    seeds = rng.integers(0, 1 << 63, size=(len(elts), k, 8), dtype=np.uint64)
    assert lib.pirw_galois_keys_save(N, MODS.ctypes.data_as(u64p), 3, C.c_uint64(T), elts.ctypes.data_as(u32p),
                                     len(elts), klimbs.ctypes.data_as(u64p), seeds.ctypes.data_as(u64p), C.byref(out),
                                     C.byref(n)) == 0
    key_blob = take(lib, out, n)
    assert lib.pirw_encryption_parameters_save(N, MODS.ctypes.data_as(u64p), 3, C.c_uint64(T), C.byref(out),
                                               C.byref(n)) == 0
    ep_blob = take(lib, out, n)

    got_ct = np.zeros_like(limbs)
    pid2 = np.zeros(4, dtype=np.uint64)
    ntt, seeded = C.c_int(), C.c_int()
    got_e = np.zeros(8, dtype=np.uint32)
    got_k = np.zeros((8, k, 2, 3, N), dtype=np.uint64)
    cnt = C.c_uint32()
    gn, gm, gt = C.c_uint32(), C.c_uint32(), C.c_uint64()
    mods = np.zeros(8, dtype=np.uint64)

    def mutate(blob):
        b = bytearray(blob)
        kind = rng.integers(0, 4)
        if kind == 0:      # flip bytes, biased to the structural first 200 bytes
            for _ in range(int(rng.integers(1, 6))):
                pos = int(rng.integers(0, min(len(b), 200))) if rng.random() < 0.7 else int(rng.integers(0, len(b)))
                b[pos] = int(rng.integers(0, 256))
        elif kind == 1:    # truncate
            b = b[:int(rng.integers(0, len(b)))]
        elif kind == 2:    # overwrite a length / count field with an extreme value
            extremes = [0, 1, 2**31, 2**32 - 1, 2**63, 2**64 - 1]
            pos = int(rng.integers(0, max(1, min(len(b) - 8, 160))))
            b[pos:pos + 8] = struct.pack("<Q", extremes[int(rng.integers(0, len(extremes)))])
        else:              # splice another object's bytes into the middle
            pos = int(rng.integers(0, len(b)))
            b = b[:pos] + bytearray(ep_blob) + b[pos:]
        return bytes(b)

    for _ in range(300):
        bad = mutate(ct_blob)
        assert lib.pirw_ct_load(buf(bad), C.c_size_t(len(bad)), N, MODS.ctypes.data_as(u64p), k,
                                got_ct.ctypes.data_as(u64p), pid2.ctypes.data_as(u64p), C.byref(ntt),
                                C.byref(seeded)) in (0, 3)
    for _ in range(60):
        bad = mutate(key_blob)
        assert lib.pirw_galois_keys_load(buf(bad), C.c_size_t(len(bad)), N, MODS.ctypes.data_as(u64p), 3,
                                         C.c_uint64(T), 8, got_e.ctypes.data_as(u32p), got_k.ctypes.data_as(u64p),
                                         C.byref(cnt)) in (0, 3)
    for _ in range(300):
        bad = mutate(ep_blob)
        assert lib.pirw_encryption_parameters_load(buf(bad), C.c_size_t(len(bad)), C.byref(gn),
                                                   mods.ctypes.data_as(u64p), 8, C.byref(gm), C.byref(gt)) in (0, 3)
    for kind, good in [(1, b"\x0a\x05\x0a\x03abc\x12\x01k"), (3, b"\x08\x05\x12\x02\x29\x28\x20\x07")]:
        for _ in range(300):
            bad = mutate(good * 3)
            rc, _ = roundtrip(lib, kind, bad)
            assert rc in (0, 3)


def test_wire_end_to_end_against_the_oracle_cpp():
    """tests/cpp/wire_test.cpp: protobuf-framed request with seed-compressed keys -> raw limbs -> oracle answer ->
    serialized reply -> decrypt.  No GPU involved; the oracle is the checker."""
    exe = os.path.join(ROOT, "build", "wire_test")
    subprocess.check_call(["make", "-C", ROOT, "build/wire_test"], stdout=subprocess.DEVNULL)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "WIRE_TEST_OK" in out.stdout, out.stdout + out.stderr


def test_wire_parsers_survive_malformed_requests_under_sanitizers():
    """tests/cpp/wire_fuzz_test.cpp (AddressSanitizer + UBSan): truncated, bit-flipped, length-inflated and spliced
    requests are rejected or parsed without out-of-bounds accesses, and the copy-free parsers of the device server's
    wire path (ParseView, LoadCiphertextTo) agree with the object-building ones on every input."""
    exe = os.path.join(ROOT, "build", "wire_fuzz_test")
    subprocess.check_call(["make", "-C", ROOT, "build/wire_fuzz_test"], stdout=subprocess.DEVNULL)
    out = subprocess.run([exe, "1500"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "WIRE_FUZZ_TEST_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


# ---------------------------------------------------------------------------------------------------------------
# pir_b200/wire.py: the Python mirror's view of the same codec (host-side only, no GPU)
# ---------------------------------------------------------------------------------------------------------------
def test_python_wire_request_response_roundtrip(pb):
    import pir_b200 as pbm
    from pir_b200 import wire
    ep = pbm.GenerateEncryptionParams(4096, 20)
    p = pbm.CreatePIRParameters(5000, 0, 1, ep)          # 2 query ciphertexts per query
    k, n = len(ep.coeff_modulus) - 1, ep.poly_modulus_degree
    rng = np.random.default_rng(21)
    mods = [int(q) for q in ep.coeff_modulus]
    queries = [np.ascontiguousarray(np.stack([rng.integers(0, mods[j], (2, 2, n), dtype=np.uint64) for j in range(k)],
                                             axis=2)) for _ in range(3)]   # [n_ct=2][2][k][N] each
    elts = pbm.generate_galois_elts(n)
    keys = np.ascontiguousarray(np.stack([rng.integers(0, mods[j], (len(elts), k, 2, n), dtype=np.uint64)
                                          for j in range(k + 1)], axis=3))
    gk = pbm.GaloisKeys(elts, keys.reshape(-1))
    relin = wire.save_galois_keys(pbm.GaloisKeys([3], keys[:1].reshape(-1)), ep)
    blob = wire.serialize_request(queries, gk, p, relin_keys=relin)
    # the protobuf runtime reads the same structure out of it
    ref = pb["Request"]()
    ref.ParseFromString(blob)
    assert len(ref.query) == 3 and all(len(q.ct) == 2 for q in ref.query)
    assert ref.relin_keys == relin and len(ref.galois_keys) > len(elts) * k * 2 * (k + 1) * n * 8
    req = wire.parse_request(blob, p)
    assert len(req.query) == 3
    for a, b in zip(req.query, queries):
        assert np.array_equal(a, b)
    order = np.argsort(elts)
    assert req.galois_keys.elts == [int(elts[i]) for i in order]
    assert np.array_equal(req.galois_keys.data.reshape(len(elts), -1), keys[order].reshape(len(elts), -1))
    assert np.array_equal(req.parms_id, wire.data_parms_id(ep))
    # responses
    resp = pbm.Response([np.ascontiguousarray(np.stack([rng.integers(0, mods[j], (1, 2, n), dtype=np.uint64)
                                                        for j in range(k)], axis=2)) for _ in range(3)])
    rb = wire.serialize_response(resp, p, parms_id=req.parms_id)
    ref_r = pb["Response"]()
    ref_r.ParseFromString(rb)
    assert len(ref_r.reply) == 3 and all(len(r.ct) == 1 for r in ref_r.reply)
    back = wire.parse_response(rb, p)
    for a, b in zip(back.reply, resp.reply):
        assert np.array_equal(a, b)
    # malformed input surfaces as the reference's InvalidArgument
    for bad in [blob[:len(blob) // 3], b"\x12\x03abc", rb]:
        with pytest.raises(pbm.PIRStatusError) as e:
            wire.parse_request(bad, p)
        assert e.value.code == pbm.INVALID_ARGUMENT


def test_python_wire_pir_parameters(pb):
    import pir_b200 as pbm
    from pir_b200 import wire
    ep = pbm.GenerateEncryptionParams(4096, 24)
    p = pbm.CreatePIRParameters(1 << 16, 288, 2, ep)
    blob = wire.serialize_pir_parameters(p)
    ref = pb["PIRParameters"]()
    ref.ParseFromString(blob)
    assert (ref.num_items, ref.num_pt, list(ref.dimensions), ref.bytes_per_item, ref.items_per_plaintext) == \
        (1 << 16, 1639, [41, 40], 288, p.items_per_plaintext)
    ep2 = wire.load_encryption_parameters(ref.encryption_parameters)
    assert (ep2.poly_modulus_degree, ep2.plain_modulus, list(ep2.coeff_modulus)) == \
        (4096, ep.plain_modulus, list(ep.coeff_modulus))


def test_process_request_bytes_composition_with_the_oracle_as_engine():
    """PIRServer.ProcessRequestBytes = parse -> ProcessRequest -> serialize.  Here ProcessRequest is played by the CPU
    oracle (a stand-in object; the GPU tests cover the real one), so the wire composition is checked end to end on the
    CPU: a harness client's query travels as bytes, the reply comes back as bytes and decrypts to the right item."""
    import pir_b200 as pbm
    from oracle import client as oc
    from pir_b200 import wire
    ep = pbm.GenerateEncryptionParams(4096, 20)
    p = pbm.CreatePIRParameters(10, 0, 1, ep)
    hp = oc.PIRParameters(p.num_items, p.num_pt, list(p.dimensions), p.bytes_per_item, p.items_per_plaintext,
                          p.bits_per_coeff, ep.poly_modulus_degree, ep.plain_modulus, list(ep.coeff_modulus))
    cl = oc.HarnessClient(hp, seed=4)
    rng = np.random.default_rng(9)
    items = [rng.integers(0, 256, p.bytes_per_item, dtype=np.uint8).tobytes() for _ in range(p.num_items)]
    db_ntt = oc.db_to_ntt(cl.orc, oc.encode_string_db(hp, items))

    class OracleServer:
        params = p

        def ProcessRequest(self, request):
            gk = request.galois_keys
            return pbm.Response([cl.orc.process_query(db_ntt, p.dimensions, gk.elts, gk.data, q)
                                 for q in request.query])

    idxs = [3, 8]
    blob = wire.serialize_request([cl.create_query(i) for i in idxs], pbm.GaloisKeys(cl.elts, cl.galois), p)
    reply_bytes = pbm.PIRServer.ProcessRequestBytes(OracleServer(), blob)
    resp = wire.parse_response(reply_bytes, p)
    assert len(resp.reply) == 2
    got = cl.process_response_strings(idxs, resp.reply)
    assert got == [items[i] for i in idxs]


def test_process_request_bytes_in_ciphertext_multiplication_mode():
    """The same composition with use_ciphertext_multiplication: the request's relin_keys field (a KSwitchKeys object
    with one slot, serialization.cpp:66-72) is deserialized into raw limbs and USED (server.cpp:53-58, 185-190); every
    reply is ONE ciphertext — 2 polynomials with keys, 3 without (d = 2) — and survives the wire round trip.  The
    oracle plays ProcessRequest; the GPU suite covers the real one."""
    import pir_b200 as pbm
    from oracle import client as oc
    from pir_b200 import wire
    ep = pbm.GenerateEncryptionParams(4096, 16)
    p = pbm.CreatePIRParameters(9, 0, 2, ep, True, 10)      # correctness_test.cpp:95
    assert p.use_ciphertext_multiplication
    hp = oc.PIRParameters(p.num_items, p.num_pt, list(p.dimensions), p.bytes_per_item, p.items_per_plaintext,
                          p.bits_per_coeff, ep.poly_modulus_degree, ep.plain_modulus, list(ep.coeff_modulus), True)
    cl = oc.HarnessClient(hp, seed=6)
    rng = np.random.default_rng(10)
    items = [rng.integers(0, 256, p.bytes_per_item, dtype=np.uint8).tobytes() for _ in range(p.num_items)]
    db_ntt = oc.db_to_ntt(cl.orc, oc.encode_string_db(hp, items))
    seen = []

    class OracleServer:
        params = p

        def ProcessRequest(self, request):
            gk = request.galois_keys
            seen.append(request.relin_keys)
            return pbm.Response([cl.orc.process_query_ct(db_ntt, p.dimensions, gk.elts, gk.data, q,
                                                         request.relin_keys)[None] for q in request.query])

    idxs = [1, 5]
    queries = [cl.create_query(i) for i in idxs]
    gk = pbm.GaloisKeys(cl.elts, cl.galois)
    relin_blob = wire.save_galois_keys(pbm.GaloisKeys([1], cl.relin), ep)   # Galois element 1 <-> slot 0
    for blob, polys in ((wire.serialize_request(queries, gk, p, relin_keys=relin_blob), 2),
                        (wire.serialize_request(queries, gk, p), 3)):
        resp = wire.parse_response(pbm.PIRServer.ProcessRequestBytes(OracleServer(), blob), p)
        assert [r.shape for r in resp.reply] == [(1, polys, cl.orc.k, 4096)] * 2
        assert cl.process_response_strings(idxs, resp.reply) == [items[i] for i in idxs]
    assert np.array_equal(seen[0], cl.relin) and seen[1] is None
    # outside this mode the field is only checked for well-formedness and dropped
    p_plain = pbm.CreatePIRParameters(9, 0, 2, ep, False, 10)
    req = wire.parse_request(wire.serialize_request(queries, gk, p_plain, relin_keys=relin_blob), p_plain)
    assert req.relin_keys is None
    with pytest.raises(pbm.PIRStatusError) as e:
        wire.parse_request(wire.serialize_request(queries, gk, p, relin_keys=relin_blob[:-9]), p)
    assert e.value.code == pbm.INVALID_ARGUMENT
