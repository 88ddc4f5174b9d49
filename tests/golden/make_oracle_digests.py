#!/usr/bin/env python
"""Regenerates tests/golden/oracle_digests.json: SHA-256 of the oracle's replies (and of intermediate objects) for
fixed seeds.  The oracle defines what "bit-exact" means for the CUDA path, so its own outputs are frozen here: any
change to oracle/ that alters a limb shows up as a digest mismatch in tests/test_oracle_kat.py.

    python tests/golden/make_oracle_digests.py        # rewrites the fixture; commit the diff only if intended
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import client as oc  # noqa: E402

CASES = [  # name, items, bytes per item (0 = one plaintext per item), dimensions, N, plain bits, index
    ("n4096_d1_10items", 10, 0, 1, 4096, 20, 7),
    ("n4096_d2_82items", 82, 0, 2, 4096, 20, 42),
    ("n4096_d2_t24_300x288B", 300, 288, 2, 4096, 24, 123),
    ("n8192_d2_20items", 20, 0, 2, 8192, 20, 11),
]


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint64).tobytes()).hexdigest()


def limbs(tag, shape_prefix, mods, n):
    """Deterministic pseudo-random limbs below each modulus from SHAKE-256 (stable across numpy / libm versions)."""
    count = int(np.prod(shape_prefix)) * len(mods) * n
    raw = np.frombuffer(hashlib.shake_256(tag.encode()).digest(8 * count), dtype=np.uint64)
    a = raw.reshape(tuple(shape_prefix) + (len(mods), n)).copy()
    for j, q in enumerate(mods):
        a[..., j, :] %= np.uint64(q)
    return np.ascontiguousarray(a)


class OracleBackend:
    """The four operations whose outputs are frozen, answered by the CPU oracle."""

    def __init__(self, orc, p, db, elts, keys):
        self.orc, self.p, self.db, self.elts, self.keys = orc, p, db, elts, keys

    def substitute(self, ct, power):
        return self.orc.substitute(ct, power, self.elts, self.keys)

    def shift(self, ct, k):
        return self.orc.mul_inv_pow_x(ct, k)

    def expand(self, cts, total):
        return np.stack(self.orc.expand(cts, total, self.elts, self.keys))

    def answer(self, query):
        return self.orc.process_query(self.db, self.p.dimensions, self.elts, self.keys, query)

    def answer_ct(self, query, relin):
        """the same query in ciphertext-multiplication mode (database.cpp:202-211): ONE ciphertext [polys][k][N]"""
        return self.orc.process_query_ct(self.db, self.p.dimensions, self.elts, self.keys, query, relin)


def case_inputs(name, items, size, d, n, bits):
    """Pure integer inputs: SHAKE-derived ring elements (no noise sampling, no floating point)."""
    p = oc.create_pir_parameters(items, size, d, n, bits)
    orc = oc.HarnessClient(p, seed=1).orc                    # only used for its Oracle handle
    k, mods = orc.k, [int(q) for q in orc.moduli]
    elts = [(n >> i) + 1 for i in range(n.bit_length() - 1)]
    keys = limbs(name + "/keys", (len(elts), k, 2), mods, n)          # [n_elts][k][2][k+1][N]
    n_ct = p.dim_sum // n + 1
    query = limbs(name + "/query", (n_ct, 2), mods[:k], n)            # [n_ct][2][k][N]
    db = limbs(name + "/db", (p.num_pt,), mods[:k], n)                # NTT-form database
    return p, orc, elts, keys, query, db


def run_case(name, items, size, d, n, bits, index, make_backend=None):
    """Digests of one case.  make_backend(p, db, elts, keys_flat) -> object with substitute / shift / expand / answer /
    answer_ct; default: the oracle.  The GPU parity suite passes the CUDA path and must reproduce the same digests."""
    p, orc, elts, keys, query, db = case_inputs(name, items, size, d, n, bits)
    flat = keys.reshape(-1)
    be = make_backend(p, db, elts, flat) if make_backend else OracleBackend(orc, p, db, elts, flat)
    reply = be.answer(query)
    relin = limbs(name + "/relin", (orc.k, 2), [int(q) for q in orc.moduli], n).reshape(-1)   # [k][2][k+1][N]
    ct_relin, ct_plain = be.answer_ct(query, relin), be.answer_ct(query, None)
    return {"inputs": digest(np.concatenate([flat, query.reshape(-1), db.reshape(-1)])),
            "substitute": digest(be.substitute(query[0], elts[0])),
            "multiply_inverse_power_of_x": digest(be.shift(query[0], 5 + index)),
            "selection_vector": digest(be.expand(query, p.dim_sum)),
            "reply": digest(reply), "reply_cts": int(reply.shape[0]),
            # ciphertext-multiplication mode on the same inputs, with and without a relinearization key
            "ct_mult_reply_relin": digest(ct_relin), "ct_mult_reply": digest(ct_plain),
            "ct_mult_polys": [int(ct_relin.shape[0]), int(ct_plain.shape[0])]}


def main():
    out = {"_about": "SHA-256 of oracle outputs for fixed seeds (see make_oracle_digests.py)"}
    for name, *args in CASES:
        out[name] = run_case(name, *args)
        print(name, out[name]["reply"][:16])
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_digests.json"), "w") as f:
        json.dump(out, f, indent=1)
        f.write("\n")


if __name__ == "__main__":
    main()
