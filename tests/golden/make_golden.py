#!/usr/bin/env python3
"""Generates the committed golden fixtures under tests/golden/.

The reference (Go, and without any ECDSA / SHA-256 vector of its own — SURVEY.md §8c) cannot produce
fixtures, so these are produced by the pinned CPU oracle: verdicts are computed by BOTH oracle
restatements (OpenSSL-backed C and pure-Python big-int) and must agree before the file is written.
Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import corpus  # noqa: E402
from oracle import ecdsa_ref as ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def ecdsa(curve, n, K, seed, name):
    b = corpus.make_batch(curve, n=n, K=K, seed=seed, corrupt_rate=2)
    ok_c = oracle.verify_batch(curve, b["r"], b["s"], b["qx"], b["qy"], b["digest"])
    ok_py = np.array([ref.verify_bytes(curve, b["r"][i].tobytes(), b["s"][i].tobytes(), b["qx"][i].tobytes(), b["qy"][i].tobytes(),
                                       b["digest"][i].tobytes()) for i in range(n)], np.uint8)
    assert (ok_c == ok_py).all(), "oracle restatements disagree"
    assert 0 < ok_c.sum() < n and len(set(b["labels"].tolist())) >= 10
    np.savez_compressed(os.path.join(HERE, name), r=b["r"], s=b["s"], qx=b["qx"], qy=b["qy"], digest=b["digest"], ok=ok_c,
                        labels=b["labels"], keys=b["keys"], key_idx=b["key_idx"])
    print(name, n, "items,", int(ok_c.sum()), "accept")


def sha():
    lens = list(range(0, 70)) + [111, 112, 119, 120, 127, 128, 129, 255, 256, 257, 1000, 4095, 4096]  # every padding boundary
    off = np.zeros(len(lens) + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    g = corpus.DRBG(5)
    msgs = np.frombuffer(g.bytes(0, int(off[-1])), np.uint8).copy()
    dig = np.stack([np.frombuffer(hashlib.sha256(msgs[int(off[i]):int(off[i + 1])].tobytes()).digest(), np.uint8) for i in range(len(lens))])
    np.savez_compressed(os.path.join(HERE, "sha256_ragged.npz"), msgs=msgs, off=off, digest=dig)
    print("sha256_ragged.npz", len(lens), "messages")


def digests():
    # Proposal.Digest / CommitSignaturesDigest / quorum table as text (restated rules; see oracle/ecdsa_ref.py)
    rows = []
    for payload, header, md, seq in [(b"\x01", b"\x00", b"\x08\x01", 1), (b"\x02", b"\x01", b"\x03", 1), (b"", b"", b"", 0),
                                     (b"\xaa" * 300, b"h", b"m" * 130, 128), (b"p", b"", b"", -1), (b"p", b"q", b"r", 2**40)]:
        rows.append("proposal %s %s %s %d %s" % (payload.hex() or "-", header.hex() or "-", md.hex() or "-", seq, ref.proposal_digest(payload, header, md, seq)))
    rows.append("commitsigs %s" % ref.commit_signatures_digest([(1, b"\x04", b"\x05"), (2, b"\x04" * 70, b"")]).hex())
    for n in range(1, 33):
        q, f = ref.compute_quorum(n)
        rows.append("quorum %d %d %d" % (n, q, f))
    open(os.path.join(HERE, "digests_and_quorum.txt"), "w").write("\n".join(rows) + "\n")
    print("digests_and_quorum.txt", len(rows), "rows")


if __name__ == "__main__":
    ecdsa(oracle.P256, 256, 8, 7, "ecdsa_p256_seed7.npz")
    ecdsa(oracle.P384, 96, 4, 9, "ecdsa_p384_seed9.npz")
    sha()
    digests()
