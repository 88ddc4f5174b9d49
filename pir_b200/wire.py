"""The reference's wire format for the Python mirror: serialized pir.Request / pir.Response / pir.PIRParameters
(pir/proto/payload.proto) whose byte fields are SEAL 3.5.6 objects (pir/cpp/serialization.h:81-138).

Thin ctypes layer over pir_b200/lib/libpirb_wire.so (pir_b200/cpp/wire.hpp is the codec; see its header for what is and
is not verified against SEAL).  Host-side only: nothing here touches the GPU.

    request  = wire.parse_request(request_bytes, params)         # -> api.Request (raw limbs)
    response = server.ProcessRequest(request)
    reply    = wire.serialize_response(response, params, parms_id=request.parms_id)
or, in one call, PIRServer.ProcessRequestBytes(request_bytes).
"""
import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

from .api import (INVALID_ARGUMENT, EncryptionParameters, GaloisKeys, PIRParameters, PIRStatusError, Request,
                  Response)

_LIB = None
_u8p, _u32p, _u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libpirb_wire.so")
        if not os.path.exists(path):
            raise ImportError("libpirb_wire.so is missing: run `make wire` (or __graft_entry__.build())")
        L = C.CDLL(path)
        L.pirw_last_error.restype = C.c_char_p
        L.pirw_msg_groups.restype = C.c_uint32
        L.pirw_msg_group_size.restype = C.c_uint32
        L.pirw_msg_groups.argtypes = [C.c_void_p]
        L.pirw_msg_group_size.argtypes = [C.c_void_p, C.c_uint32]
        L.pirw_msg_free.argtypes = [C.c_void_p]
        L.pirw_msg_ct.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(_u8p), C.POINTER(C.c_size_t)]
        L.pirw_msg_keys.argtypes = [C.c_void_p, C.c_int, C.POINTER(_u8p), C.POINTER(C.c_size_t)]
        _LIB = L
    return _LIB


def _fail(what):
    raise PIRStatusError(INVALID_ARGUMENT, "%s: %s" % (what, lib().pirw_last_error().decode()))


def _buf(b: bytes):
    return (C.c_uint8 * max(1, len(b))).from_buffer_copy(b if len(b) else b"\0")


def _take(out, n) -> bytes:
    b = C.string_at(out, n.value)
    lib().pirw_free(out)
    return b


def _mods(ep: EncryptionParameters):
    return np.array([int(q) for q in ep.coeff_modulus], dtype=np.uint64)


# ----------------------------------------------------------------------------------------------------------------
# SEAL objects
# ----------------------------------------------------------------------------------------------------------------
def data_parms_id(ep: EncryptionParameters) -> np.ndarray:
    """parms_id of the data level (the special prime dropped): what query and reply ciphertexts carry."""
    m = _mods(ep)
    out = np.zeros(4, dtype=np.uint64)
    lib().pirw_parms_id(ep.poly_modulus_degree, m.ctypes.data_as(_u64p), len(m) - 1, C.c_uint64(ep.plain_modulus),
                        out.ctypes.data_as(_u64p))
    return out


def save_ciphertext(ct: np.ndarray, ep: EncryptionParameters, parms_id: Optional[np.ndarray] = None) -> bytes:
    """SEALSerialize<Ciphertext> of a data-level ciphertext [polys][k][N] (2 polynomials everywhere except replies of
    the ciphertext-multiplication mode without relinearization keys)."""
    k, n = len(ep.coeff_modulus) - 1, ep.poly_modulus_degree
    a = np.ascontiguousarray(ct, dtype=np.uint64).reshape(-1, k, n)
    pid = np.ascontiguousarray(parms_id if parms_id is not None else data_parms_id(ep), dtype=np.uint64)
    out, ln = _u8p(), C.c_size_t()
    if lib().pirw_ct_save(a.ctypes.data_as(_u64p), a.shape[0], n, k, pid.ctypes.data_as(_u64p), 0, None, C.byref(out),
                          C.byref(ln)):
        _fail("ciphertext")
    return _take(out, ln)


def load_ciphertext(blob: bytes, ep: EncryptionParameters, max_polys: int = 2):
    """SEALDeserialize<Ciphertext>: -> ([2][k][N] limbs, parms_id).  InvalidArgument on malformed input.
    max_polys > 2: accept larger ciphertexts (-> [polys][k][N])."""
    k, n = len(ep.coeff_modulus) - 1, ep.poly_modulus_degree
    m = _mods(ep)
    if max_polys > 2:
        out = np.zeros((max_polys, k, n), dtype=np.uint64)
        pid = np.zeros(4, dtype=np.uint64)
        ntt, seeded, size = C.c_int(), C.c_int(), C.c_uint32()
        if lib().pirw_ct_load_any(_buf(blob), C.c_size_t(len(blob)), n, m.ctypes.data_as(_u64p), k,
                                  out.ctypes.data_as(_u64p), max_polys, C.byref(size), pid.ctypes.data_as(_u64p),
                                  C.byref(ntt), C.byref(seeded)):
            _fail("ciphertext")
        return out[:size.value].copy(), pid
    out = np.zeros((2, k, n), dtype=np.uint64)
    pid = np.zeros(4, dtype=np.uint64)
    ntt, seeded = C.c_int(), C.c_int()
    if lib().pirw_ct_load(_buf(blob), C.c_size_t(len(blob)), n, m.ctypes.data_as(_u64p), k, out.ctypes.data_as(_u64p),
                          pid.ctypes.data_as(_u64p), C.byref(ntt), C.byref(seeded)):
        _fail("ciphertext")
    if ntt.value:
        raise PIRStatusError(INVALID_ARGUMENT, "query ciphertexts must be in coefficient form")
    return out, pid


def save_galois_keys(gk: GaloisKeys, ep: EncryptionParameters, seeds: Optional[np.ndarray] = None) -> bytes:
    """SEALSerialize<GaloisKeys>.  `seeds` ([n][k][8] u64) writes them seed-compressed: only valid when every key's
    second polynomial IS the expansion of its seed (what a SEAL client's keygen produces)."""
    m = _mods(ep)
    elts = np.array(gk.elts, dtype=np.uint32)
    data = np.ascontiguousarray(gk.data, dtype=np.uint64)
    sp = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint64).ctypes.data_as(_u64p)
    out, ln = _u8p(), C.c_size_t()
    if lib().pirw_galois_keys_save(ep.poly_modulus_degree, m.ctypes.data_as(_u64p), len(m), C.c_uint64(ep.plain_modulus),
                                   elts.ctypes.data_as(_u32p), len(elts), data.ctypes.data_as(_u64p), sp, C.byref(out),
                                   C.byref(ln)):
        _fail("galois keys")
    return _take(out, ln)


def load_galois_keys(blob: bytes, ep: EncryptionParameters) -> GaloisKeys:
    """SEALDeserialize<GaloisKeys> (seed-compressed keys are re-expanded) -> raw-limb keys in index order."""
    m = _mods(ep)
    k, n = len(m) - 1, ep.poly_modulus_degree
    max_n = 2 * n.bit_length() + 8  # the server path uses log2(N) elements; leave room for a full rotation set
    elts = np.zeros(max_n, dtype=np.uint32)
    data = np.zeros((max_n, k, 2, k + 1, n), dtype=np.uint64)
    cnt = C.c_uint32()
    if lib().pirw_galois_keys_load(_buf(blob), C.c_size_t(len(blob)), n, m.ctypes.data_as(_u64p), len(m),
                                   C.c_uint64(ep.plain_modulus), max_n, elts.ctypes.data_as(_u32p),
                                   data.ctypes.data_as(_u64p), C.byref(cnt)):
        _fail("galois keys")
    return GaloisKeys([int(e) for e in elts[:cnt.value]], data[:cnt.value].reshape(-1))


def save_encryption_parameters(ep: EncryptionParameters) -> bytes:
    m = _mods(ep)
    out, ln = _u8p(), C.c_size_t()
    if lib().pirw_encryption_parameters_save(ep.poly_modulus_degree, m.ctypes.data_as(_u64p), len(m),
                                             C.c_uint64(ep.plain_modulus), C.byref(out), C.byref(ln)):
        _fail("encryption parameters")
    return _take(out, ln)


def load_encryption_parameters(blob: bytes) -> EncryptionParameters:
    n, cnt, t = C.c_uint32(), C.c_uint32(), C.c_uint64()
    mods = np.zeros(16, dtype=np.uint64)
    if lib().pirw_encryption_parameters_load(_buf(blob), C.c_size_t(len(blob)), C.byref(n), mods.ctypes.data_as(_u64p),
                                             16, C.byref(cnt), C.byref(t)):
        _fail("encryption parameters")
    return EncryptionParameters(int(n.value), int(t.value), [int(x) for x in mods[:cnt.value]])


# ----------------------------------------------------------------------------------------------------------------
# payload.proto messages
# ----------------------------------------------------------------------------------------------------------------
def serialize_pir_parameters(p: PIRParameters) -> bytes:
    dims = np.array(list(p.dimensions), dtype=np.uint32)
    epb = save_encryption_parameters(p.encryption_parameters)
    out, ln = _u8p(), C.c_size_t()
    if lib().pirw_params_build(C.c_uint64(p.num_items), C.c_uint64(p.num_pt), dims.ctypes.data_as(_u32p), len(dims),
                               _buf(epb), C.c_size_t(len(epb)), int(p.bytes_per_item), int(p.items_per_plaintext),
                               int(p.bits_per_coeff), int(bool(p.use_ciphertext_multiplication)), C.byref(out),
                               C.byref(ln)):
        _fail("PIRParameters")
    return _take(out, ln)


class _Msg:
    def __init__(self, kind: int, data: bytes):
        self.h = C.c_void_p()
        if lib().pirw_msg_parse(kind, _buf(data), C.c_size_t(len(data)), C.byref(self.h)):
            _fail("Request" if kind == 1 else "Response")

    def groups(self) -> List[List[bytes]]:
        L = lib()
        out = []
        for g in range(L.pirw_msg_groups(self.h)):
            row = []
            for i in range(L.pirw_msg_group_size(self.h, g)):
                p, n = _u8p(), C.c_size_t()
                L.pirw_msg_ct(self.h, g, i, C.byref(p), C.byref(n))
                row.append(C.string_at(p, n.value))
            out.append(row)
        return out

    def keys(self, which: int) -> bytes:
        p, n = _u8p(), C.c_size_t()
        lib().pirw_msg_keys(self.h, which, C.byref(p), C.byref(n))
        return C.string_at(p, n.value) if n.value else b""

    def close(self):
        if self.h:
            lib().pirw_msg_free(self.h)
            self.h = None


def parse_request(data: bytes, params: PIRParameters) -> Request:
    """serialized pir.Request -> api.Request of raw limbs (server.cpp:44-58).  The relinearization keys are checked for
    well-formedness and dropped (unused on the re-encoder path).  The request's parms_id is kept in `.parms_id`."""
    ep = params.encryption_parameters
    msg = _Msg(1, data)
    try:
        gk = load_galois_keys(msg.keys(2), ep)
        rk = msg.keys(3)
        relin = None
        if rk and params.use_ciphertext_multiplication:
            # server.cpp:53-58, 185-190: RelinKeys is a KSwitchKeys object whose slot 0 holds the one key
            rkeys = load_galois_keys(rk, ep)
            if len(rkeys.elts) != 1:
                raise PIRStatusError(INVALID_ARGUMENT, "relinearization keys must hold exactly one key")
            relin = rkeys.data
        elif rk:
            m = _mods(ep)
            if lib().pirw_kswitch_keys_check(_buf(rk), C.c_size_t(len(rk)), ep.poly_modulus_degree,
                                             m.ctypes.data_as(_u64p), len(m), C.c_uint64(ep.plain_modulus)):
                _fail("relin keys")
        req = Request(galois_keys=gk, relin_keys=relin)
        req.parms_id = None
        for group in msg.groups():
            cts = []
            for blob in group:
                ct, pid = load_ciphertext(blob, ep)
                if req.parms_id is None:
                    req.parms_id = pid
                cts.append(ct)
            k, n = len(ep.coeff_modulus) - 1, ep.poly_modulus_degree
            req.query.append(np.stack(cts) if cts else np.zeros((0, 2, k, n), dtype=np.uint64))
        return req
    finally:
        msg.close()


def serialize_request(queries: Sequence[np.ndarray], gk: GaloisKeys, params: PIRParameters,
                      relin_keys: bytes = b"", seeds: Optional[np.ndarray] = None) -> bytes:
    """Client-side counterpart (tests / harness): queries[i] is [n_ct][2][k][N]."""
    ep = params.encryption_parameters
    blobs = [[save_ciphertext(ct, ep) for ct in q] for q in queries]
    per = len(blobs[0]) if blobs else 0
    if any(len(b) != per for b in blobs):
        raise PIRStatusError(INVALID_ARGUMENT, "every query must have the same number of ciphertexts")
    flat = b"".join(b"".join(b) for b in blobs)
    ct_len = len(blobs[0][0]) if per else 0
    gkb = save_galois_keys(gk, ep, seeds)
    out, ln = _u8p(), C.c_size_t()
    if lib().pirw_request_build(_buf(flat), len(blobs), per, C.c_size_t(ct_len), _buf(gkb), C.c_size_t(len(gkb)),
                                _buf(relin_keys), C.c_size_t(len(relin_keys)), C.byref(out), C.byref(ln)):
        _fail("Request")
    return _take(out, ln)


def serialize_response(resp: Response, params: PIRParameters, parms_id: Optional[np.ndarray] = None) -> bytes:
    """api.Response -> serialized pir.Response; reply ciphertexts carry `parms_id` (default: the data level's)."""
    ep = params.encryption_parameters
    blobs = [[save_ciphertext(ct, ep, parms_id) for ct in r] for r in resp.reply]
    per = len(blobs[0]) if blobs else 0
    flat = b"".join(b"".join(b) for b in blobs)
    ct_len = len(blobs[0][0]) if per else 0
    out, ln = _u8p(), C.c_size_t()
    if lib().pirw_response_build(_buf(flat), len(blobs), per, C.c_size_t(ct_len), C.byref(out), C.byref(ln)):
        _fail("Response")
    return _take(out, ln)


def parse_response(data: bytes, params: PIRParameters) -> Response:
    ep = params.encryption_parameters
    msg = _Msg(2, data)
    try:
        resp = Response()
        cap = len(params.dimensions) + 1 if params.use_ciphertext_multiplication else 2
        for group in msg.groups():
            resp.reply.append(np.stack([load_ciphertext(b, ep, cap)[0] for b in group]))
        return resp
    finally:
        msg.close()
