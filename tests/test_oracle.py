"""Pins the CPU oracle (oracle/) — the reference itself holds no vector for this path,
so the oracle is pinned on RFC 6979 A.2.5 / A.2.6 known answers, FIPS 180-4 SHA-256, OpenSSL's own
ECDSA_do_verify, python `cryptography`, and the two independent restatements against each other."""
import hashlib

import numpy as np
import pytest

import oracle
from oracle import P256, P384, corpus
from oracle import ecdsa_ref as ref

H = bytes.fromhex
from vectors import RFC6979  # noqa: E402


def _arr(hexstr):
    return np.frombuffer(H(hexstr), np.uint8)[None, :]


@pytest.mark.parametrize("vec", RFC6979)
def test_rfc6979_known_answers(vec):
    curve, ux, uy, msg, r, s = vec
    dig = hashlib.sha256(msg).digest()
    assert ref.verify(curve, int(ux, 16), int(uy, 16), dig, int(r, 16), int(s, 16))
    ok = oracle.verify_batch(curve, _arr(r), _arr(s), _arr(ux), _arr(uy), np.frombuffer(dig, np.uint8)[None, :])
    assert ok.tolist() == [1]
    # any single-bit change of r rejects
    r2 = bytearray(H(r)); r2[5] ^= 1
    assert not ref.verify(curve, int(ux, 16), int(uy, 16), dig, int.from_bytes(r2, "big"), int(s, 16))
    ok = oracle.verify_batch(curve, np.frombuffer(bytes(r2), np.uint8)[None, :], _arr(s), _arr(ux), _arr(uy),
                             np.frombuffer(dig, np.uint8)[None, :])
    assert ok.tolist() == [0]


def test_sha256_fips_vectors():
    msgs = [b"abc", b"", b"abcdbcdecdefdefgefghfghighijhijkijkljklmklmnlmnomnopnopq"]
    want = ["ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad",
            "e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855",
            "248d6a61d20638b8e5c026930c3e6039a33ce45964ff2167f6ecedd419db06c1"]
    off = np.cumsum([0] + [len(m) for m in msgs]).astype(np.uint64)
    out = oracle.sha256_batch(np.frombuffer(b"".join(msgs), np.uint8), off)
    assert [bytes(o).hex() for o in out] == want


def test_c_oracle_matches_python_restatement_and_openssl_on_corrupted_corpus():
    b = corpus.make_batch(P256, n=192, K=8, seed=11, corrupt_rate=2)
    ok = oracle.verify_batch(P256, b["r"], b["s"], b["qx"], b["qy"], b["digest"])
    py = [ref.verify_bytes(P256, b["r"][i].tobytes(), b["s"][i].tobytes(), b["qx"][i].tobytes(),
                           b["qy"][i].tobytes(), b["digest"][i].tobytes()) for i in range(b["n"])]
    assert ok.tolist() == [int(v) for v in py]
    lab = b["labels"]
    assert ok[lab == -1].all()                 # untouched items accept
    assert ok[lab == 11].all()                 # high-S still accepts (no low-S rule)
    assert not ok[(lab >= 0) & (lab != 11)].any()
    assert len(set(lab.tolist())) >= 10        # the corpus really covers the classes
    # OpenSSL's own production verifier agrees wherever the key is a valid key
    valid_key = np.array([ref.on_curve(ref.CURVES[P256], int.from_bytes(b["qx"][i].tobytes(), "big"),
                                       int.from_bytes(b["qy"][i].tobytes(), "big")) for i in range(b["n"])])
    keys = np.concatenate([b["qx"], b["qy"]], axis=1)[valid_key]
    idx = np.arange(keys.shape[0], dtype=np.uint32)
    _, ok2 = oracle.bench_verify(P256, b["r"][valid_key], b["s"][valid_key], keys, idx, b["digest"][valid_key], nthreads=2)
    assert ok2.tolist() == ok[valid_key].tolist()


def test_p384_restatements_agree():
    b = corpus.make_batch(P384, n=48, K=4, seed=21, corrupt_rate=2)
    ok = oracle.verify_batch(P384, b["r"], b["s"], b["qx"], b["qy"], b["digest"])
    py = [ref.verify_bytes(P384, b["r"][i].tobytes(), b["s"][i].tobytes(), b["qx"][i].tobytes(),
                           b["qy"][i].tobytes(), b["digest"][i].tobytes()) for i in range(b["n"])]
    assert ok.tolist() == [int(v) for v in py]
    assert ok[b["labels"] == -1].all()


def test_python_cryptography_cross_check():
    crypto = pytest.importorskip("cryptography")
    from cryptography.exceptions import InvalidSignature
    from cryptography.hazmat.primitives import hashes
    from cryptography.hazmat.primitives.asymmetric import ec, utils

    b = corpus.make_batch(P256, n=64, K=4, seed=31, corrupt_rate=3)
    ok = oracle.verify_batch(P256, b["r"], b["s"], b["qx"], b["qy"], b["digest"])
    for i in range(b["n"]):
        f = lambda a: int.from_bytes(a[i].tobytes(), "big")
        try:
            pk = ec.EllipticCurvePublicNumbers(f(b["qx"]), f(b["qy"]), ec.SECP256R1()).public_key()
        except ValueError:
            assert ok[i] == 0
            continue
        sig = ref.der_encode(f(b["r"]), f(b["s"]))
        try:
            pk.verify(sig, b["digest"][i].tobytes(), ec.ECDSA(utils.Prehashed(hashes.SHA256())))
            good = 1
        except InvalidSignature:
            good = 0
        assert good == ok[i], (i, b["labels"][i])


def test_der_strictness():
    r, s = 0x1234, (1 << 255) | 5
    good = ref.der_encode(r, s)
    assert ref.der_parse(good) == (r, s)
    bad = [
        good + b"\x00",                                    # trailing byte
        good[:-1],                                         # truncated
        b"\x30\x06\x02\x01\x01\x02\x01",                   # length mismatch
        b"\x30\x07\x02\x02\x00\x01\x02\x01\x01",           # non-minimal integer
        b"\x30\x06\x02\x01\x81\x02\x01\x01",               # negative integer
        b"\x30\x81\x06\x02\x01\x01\x02\x01\x01",           # non-minimal length form
        b"\x31\x06\x02\x01\x01\x02\x01\x01",               # wrong outer tag
        b"\x30\x06\x02\x00\x02\x02\x00\x01",               # empty integer
        b"",
    ]
    for sg in bad:
        assert ref.der_parse(sg) is None, sg.hex()
    L = 32
    import ctypes as C
    for sg in [good] + bad:
        rb = (C.c_uint8 * L)(); sb = (C.c_uint8 * L)()
        buf = (C.c_uint8 * max(len(sg), 1)).from_buffer_copy(sg or b"\x00")
        rc = oracle.lib().orc_der_parse(buf, C.c_size_t(len(sg)), C.c_size_t(L), rb, sb)
        assert rc == (1 if sg is good else 0), sg.hex()
        if sg is good:
            assert int.from_bytes(bytes(rb), "big") == r and int.from_bytes(bytes(sb), "big") == s


def test_quorum_golden_table():
    # /root/reference/internal/bft/util_test.go:144-154 (N, F, Q)
    table = [(4, 1, 3), (5, 1, 4), (6, 1, 4), (7, 2, 5), (8, 2, 6), (9, 2, 6), (10, 3, 7), (11, 3, 8), (12, 3, 8)]
    for n, f, q in table:
        assert ref.compute_quorum(n) == (q, f)
    assert ref.compute_quorum(16) == (11, 5)


def test_proposal_digest_fixtures():
    # DER derived from the fixtures at internal/bft/view_test.go:32-49 (values derived in SURVEY §8c,
    # not present in the reference): checks the DER framing + SHA-256 path end to end.
    assert ref.proposal_der(b"\x01", b"\x00", b"\x08\x01", 1).hex() == "300d04010104010004020801020101"
    assert ref.proposal_digest(b"\x01", b"\x00", b"\x08\x01", 1) == hashlib.sha256(H("300d04010104010004020801020101")).hexdigest()
    assert ref.proposal_der(b"\x02", b"\x01", b"\x03", 1).hex() == "300c040102040101040103020101"
    # long-form lengths and negative / multi-byte INTEGER
    big = ref.proposal_der(b"\xaa" * 300, b"", b"", 128)
    assert big[:4].hex() == "30820138" and big[4:8].hex() == "0482012c"
    assert ref._der_int64(128).hex() == "02020080" and ref._der_int64(-1).hex() == "0201ff"


def test_vote_counting_rules():
    # TestValidateLastDecision "not enough valid signatures": signers 0,0,1 all verifying, quorum 3
    assert ref.validate_last_decision_sigs([0, 0, 1], [1, 1, 1], 3) is False
    assert ref.validate_last_decision_sigs([1, 2, 3], [1, 1, 1], 3) is True
    assert ref.validate_last_decision_sigs([1, 2, 3], [1, 0, 1], 3) is False
    assert ref.validate_last_decision_sigs([1, 2], [1, 1], 3) is False
    # TestNormalPath: two foreign commits reach Quorum-1 = 2; TestBadCommit: a bad vote is dropped
    assert ref.count_commit_votes([(2, 2, 1, 1), (3, 3, 1, 1)], self_id=1) == 2
    assert ref.count_commit_votes([(2, 2, 0, 1), (3, 3, 1, 0), (4, 4, 1, 1)], self_id=1) == 1
    # double vote: the first registered vote burns the sender's slot
    assert ref.count_commit_votes([(2, 2, 1, 0), (2, 2, 1, 1)], self_id=1) == 0
    # signer != sender is not registered, so the sender may still vote properly later
    assert ref.count_commit_votes([(2, 3, 1, 1), (2, 2, 1, 1)], self_id=1) == 1
