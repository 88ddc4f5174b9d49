"""CPU-only checks of the product's host side: the C-ABI library loads and exports every declared symbol,
and the host logic (shape math, parameters, string packing) matches the reference's known answers and the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import pir_b200 as pb
from pir_b200 import _lib
from oracle import binding as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "pir_b200.h")).read()
    declared = set(re.findall(r"\b(pirb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = C.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), "library does not export %s" % name
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    _lib.lib()  # binds argtypes for all of them


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    p = pb.CreatePIRParameters(10, 0, 1)
    with pytest.raises(pb.PIRStatusError) as e:
        pb.PIRDatabase.Create(p)
    assert e.value.code == pb.INTERNAL and "no CUDA device" in e.value.message


def test_shape_math_matches_reference_tables():
    # utils_test.cpp:24-63
    for v, e in [(0, 1), (1, 1), (2, 2), (3, 4), (8, 8), (9, 16), ((1 << 16) + 1, 131072), ((1 << 30) + 1, 1 << 31)]:
        assert pb.next_power_two(v) == e
    for v, e in [(1, 0), (2, 1), (3, 2), (8, 3), (15, 4), (16, 4), (17, 5), ((1 << 16) + 1, 17), (1 << 31, 31)]:
        assert pb.ceil_log2(v) == e
    for v, e in [(1, 0), (2, 1), (3, 1), (8, 3), (15, 3), (16, 4), ((1 << 16) - 1, 15), ((1 << 31) - 1, 30)]:
        assert pb.log2(v) == e
    # database_test.cpp:456-464
    for n, d, e in [(100, 1, [100]), (100, 2, [10, 10]), (82, 2, [10, 9]), (975, 2, [32, 31]),
                    (1000, 3, [10, 10, 10]), (1001, 3, [11, 10, 10]), (1000001, 3, [101, 100, 100])]:
        assert pb.calculate_dimensions(n, d) == e
    assert pb.generate_galois_elts(4096) == ob.generate_galois_elts(4096)
    rng = np.random.default_rng(0)
    for _ in range(200):
        n, d = int(rng.integers(1, 1 << 24)), int(rng.integers(1, 5))
        assert pb.calculate_dimensions(n, d) == ob.calculate_dimensions(n, d)


def test_encryption_params_match_oracle():
    for n, bits in [(4096, 20), (4096, 16), (4096, 24), (8192, 20), (4096, 22)]:
        ep = pb.GenerateEncryptionParams(n, bits)
        assert ep.plain_modulus == ob.plain_modulus_batching(n, bits)
        assert ep.coeff_modulus == ob.bfv_default(n)
    assert pb.GenerateEncryptionParams(4096, 20).plain_modulus == 0xFC001  # server_test.cpp:295


def test_create_pir_parameters_kat():
    # parameters_test.cpp:49-70
    p = pb.CreatePIRParameters(1026, 256)
    assert (p.num_items, p.num_pt, p.bytes_per_item, p.items_per_plaintext, p.dimensions) == (1026, 27, 256, 38, [27])
    p = pb.CreatePIRParameters(19011, 500, 3)
    assert (p.num_pt, p.items_per_plaintext, p.dimensions) == (1001, 19, [11, 10, 10])
    # parameters_test.cpp:81-91 (shape part; CT-multiplication itself is out of scope)
    p = pb.CreatePIRParameters(77412, 777, 2, pb.GenerateEncryptionParams(8192), False, 12)
    assert (p.num_pt, p.items_per_plaintext, p.dimensions, p.bits_per_coeff) == (5161, 15, [72, 72], 12)
    with pytest.raises(pb.PIRStatusError):
        pb.CreatePIRParameters(10, 9729, 1)  # cannot fit an item
    with pytest.raises(pb.PIRStatusError):
        pb.CreatePIRParameters(10, 64, 1, None, False, 21)  # bits per coeff > max
    # BASELINE.json configs
    p = pb.CreatePIRParameters(1 << 16, 288, 2, pb.GenerateEncryptionParams(4096, 24))
    assert (p.num_pt, p.dimensions) == (1639, [41, 40])
    p = pb.CreatePIRParameters(1 << 22, 256, 2)
    assert (p.num_pt, p.dimensions) == (110377, [333, 332])
    p = pb.CreatePIRParameters(1 << 20, 1024, 2, pb.GenerateEncryptionParams(8192, 20))
    assert (p.num_pt, p.dimensions) == (55189, [235, 235])


def test_string_encoder_matches_oracle_and_reference_shapes():
    ep = pb.GenerateEncryptionParams(4096, 20)
    enc = pb.StringEncoder(ep)
    # string_encoder_test.cpp:64-71
    assert [enc.num_items_per_plaintext(s) for s in (1, 9728, 9729, 99999, 64, 288)] == [9728, 1, 0, 0, 152, 33]
    assert enc.max_bytes_per_plaintext() == 9728
    rng = np.random.default_rng(1)
    for bits in (19, 15, 10, 6, 23):
        enc.set_bits_per_coeff(bits)
        for n in (1, 7, 64, 289, 1000, 4096 * bits // 8):
            blob = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
            c = enc.encode(blob)
            assert np.array_equal(c, ob.string_encode(blob, bits, 4096))
            assert enc.decode(c, n) == blob
            off = n // 3
            assert enc.decode(c, n - off, off) == ob.string_decode(c, bits, n - off, off)
    enc.set_bits_per_coeff(19)
    with pytest.raises(pb.PIRStatusError):
        enc.encode(b"x" * 9729)
    value = b"This is a string test for random VALUES@!#"  # string_encoder_test.cpp:73-83
    assert enc.decode(enc.encode(value))[:len(value)] == value


def test_boundary_headers_compile_strictly(tmp_path):
    """include/pir_b200.h is plain C99 (the C ABI a foreign-function binding would parse); the C++ shim and the wire
    codec compile warning-free with -Wall -Wextra -pedantic."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    c = tmp_path / "t.c"
    c.write_text('#include "include/pir_b200.h"\nint main(void) { return 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", root, str(c)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cpp = tmp_path / "t.cpp"
    cpp.write_text('#include "pir_b200/cpp/pir_b200.hpp"\n#include "pir_b200/cpp/wire.hpp"\nint main() { return 0; }\n')
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", root,
                        str(cpp)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_process_request_marshalling_with_a_stub_library(monkeypatch):
    """Host-side argument marshalling of PIRServer.ProcessRequest without a device: the C entry point is replaced by a
    stub that records what it is given (addresses of the caller's own buffers, counts) — no compute involved."""
    import types
    from pir_b200 import api
    ep = pb.GenerateEncryptionParams(4096, 24)
    p = pb.CreatePIRParameters(1 << 16, 288, 2, ep)
    srv = api.PIRServer.__new__(api.PIRServer)
    srv.params = p
    srv.ctx = types.SimpleNamespace(ct_limbs=2 * 2 * 4096, query_cts=1, reply_cts=8, k=2, N=4096, h=None)
    gk = pb.GaloisKeys([3], np.zeros(8, dtype=np.uint64))
    srv._key_cache = (gk, types.SimpleNamespace(h=None))
    seen = {}

    class Stub:
        def pirb_answer(self, h, kh, qp, nq, nct, op):
            seen.update(q=qp.value, o=op.value, n=(nq, nct))
            return 0

    monkeypatch.setattr(api._lib, "lib", lambda: Stub())
    q = np.zeros((1, 2, 2, 4096), dtype=np.uint64)
    out = np.zeros((2, 8, 2, 2, 4096), dtype=np.uint64)
    r = srv.ProcessRequest(pb.Request([q, q.copy()], gk), out=out)
    assert seen["o"] == out.ctypes.data and seen["n"] == (2, 1)
    assert [x.shape for x in r.reply] == [(8, 2, 2, 4096)] * 2 and r.reply[1].ctypes.data == out[1].ctypes.data
    r = srv.ProcessRequest(pb.Request([q], gk), out=out[:1])
    assert seen["q"] == q.ctypes.data and seen["n"] == (1, 1)          # a single contiguous query is not copied
    assert srv.ProcessRequest(pb.Request([q], gk)).reply[0].shape == (8, 2, 2, 4096)
    with pytest.raises(pb.PIRStatusError) as e:
        srv.ProcessRequest(pb.Request([q], gk), out=np.zeros(5, dtype=np.uint64))
    assert e.value.code == pb.INVALID_ARGUMENT
    with pytest.raises(pb.PIRStatusError):                             # wrong number of query ciphertexts
        srv.ProcessRequest(pb.Request([np.zeros((2, 2, 2, 4096), dtype=np.uint64)], gk))
    assert srv.ProcessRequest(pb.Request([], gk)).reply == []


def test_cpp_shim_host_side_tables():
    """tests/cpp/shim_host_test.cpp: the reference's parameters / string-encoder / index tables through the C++ shim."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-C", root, "build/shim_host_test"], stdout=subprocess.DEVNULL)
    out = subprocess.run([os.path.join(root, "build", "shim_host_test")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "SHIM_HOST_TEST_OK" in out.stdout, out.stdout + out.stderr
