"""BASELINE.json configs at FULL size on the GPU (C2: 64K P-256; C3: 1M requests hash+verify; C4: n=16
f=5 quorum stream, 262,144 votes; C5: 64K mixed P-256/P-384).  The multi-threaded oracle is fast
enough for exact comparison at these sizes; on top of that the tests check size-independent
properties (tiling invariance, corruption => reject, quorum monotonicity)."""
import numpy as np
import pytest

import oracle
from oracle import P256, P384, corpus
from oracle import ecdsa_ref as ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import consensus_b200 as sbv
    e = sbv.Engine(n_devices=1)
    yield e
    e.close()


@pytest.fixture(scope="module")
def c2():
    return corpus.make_batch(P256, n=65536, K=1024, seed=1)


def test_c2_64k_p256_exact(eng, c2):
    b = c2
    want = oracle.verify_batch(P256, b["r"], b["s"], b["qx"], b["qy"], b["digest"])
    got = eng.verify_batch(P256, b["r"], b["s"], b["qx"], b["qy"], b["digest"])
    assert (want == got).all()
    lab = b["labels"]
    assert got[lab == -1].all() and got[lab == 11].all() and not got[(lab >= 0) & (lab != 11)].any()
    assert set(range(12)) <= set(lab.tolist())
    # idempotence and permutation invariance: verdicts follow the items
    perm = np.random.default_rng(1).permutation(b["n"])
    got_p = eng.verify_batch(P256, b["r"][perm], b["s"][perm], b["qx"][perm], b["qy"][perm], b["digest"][perm])
    assert (got_p == got[perm]).all()


def test_c3_1m_requests_hash_then_verify(eng, c2):
    """1,048,576 requests = the 65,536 signed C2 items tiled 16x with 256-byte payloads whose SHA-256
    is the signed digest... the payloads are fresh, so signatures are re-made for one tile."""
    tile, T = 65536, 16
    msgs1, off1 = corpus.make_requests(tile, seed=5, fixed_len=256)
    dig1 = oracle.sha256_batch(msgs1, off1)
    d, kxy = corpus.make_keys(P256, 4096, seed=71)
    key_idx = (np.arange(tile) % 4096).astype(np.uint32)
    r1, s1 = oracle.sign_batch(P256, d, key_idx, dig1, corpus._blocks(73, tile, 32, b"k"))
    msgs1 = msgs1.copy()
    flip = np.arange(0, tile, 16)                       # 1/16: flip one payload bit
    msgs1[(off1[flip] + (flip % 256)).astype(np.int64)] ^= 1
    s1[np.arange(5, tile, 16), 7] ^= 0x10               # 1/16: flip one signature bit
    qx1, qy1 = np.ascontiguousarray(kxy[key_idx, :32]), np.ascontiguousarray(kxy[key_idx, 32:])
    want1 = oracle.verify_batch(P256, r1, s1, qx1, qy1, oracle.sha256_batch(msgs1, off1))
    msgs = np.tile(msgs1, T)
    off = np.arange(tile * T + 1, dtype=np.uint64) * 256
    rep = lambda a: np.tile(a, (T, 1))
    got, dig = eng.hash_verify_batch(P256, msgs, off, rep(r1), rep(s1), rep(qx1), rep(qy1), want_digest=True)
    assert got.size == 1048576
    assert (dig == oracle.sha256_batch(msgs, off)).all()          # all 1M digests exact
    assert (got.reshape(T, tile) == want1[None, :]).all()         # every tile reproduces the oracle verdicts
    assert 0.8 < want1.mean() < 0.9


def test_c3_ragged_request_lengths(eng):
    n = 20000
    msgs, off = corpus.make_requests(n, seed=9, fixed_len=None, lo=64, hi=10240)   # RequestMaxBytes = 10 KiB
    assert (eng.sha256_batch(msgs, off) == oracle.sha256_batch(msgs, off)).all()


def test_c4_quorum_stream_n16(eng):
    N = 16
    q, f = ref.compute_quorum(N)
    assert (q, f) == (11, 5)
    I = 17476
    votes_per = N - 1
    n = I * votes_per                                   # 262,140 votes (+4 padding below)
    d, kxy = corpus.make_keys(P256, N, seed=81)
    rng = np.random.default_rng(6)
    inst = np.repeat(np.arange(I, dtype=np.uint32), votes_per)
    sender = np.tile(np.arange(1, N, dtype=np.uint16), I)             # self = node 0
    signer = sender.copy()
    digest_match = np.ones(n, np.uint8)
    proposal_digest = corpus.make_digests(I, seed=83)                  # one proposal digest per instance
    dig = proposal_digest[inst]
    r, s = oracle.sign_batch(P256, d, sender.astype(np.uint32), dig, corpus._blocks(85, n, 32, b"k"))
    # per instance b ~ U{0..5} Byzantine votes: bad signature / wrong digest / duplicate sender / signer != sender
    for i in range(I):
        for j in rng.choice(votes_per, size=int(rng.integers(0, 6)), replace=False):
            v = i * votes_per + int(j)
            kind = int(rng.integers(0, 4))
            if kind == 0: s[v, int(rng.integers(32))] ^= 1 << int(rng.integers(8))
            elif kind == 1: digest_match[v] = 0
            elif kind == 2 and j > 0: sender[v] = sender[v - 1]; signer[v] = signer[v - 1]   # second vote of a sender
            elif kind == 3: signer[v] = (signer[v] % (N - 1)) + 1 if signer[v] != 1 else 2
    # 4 padding votes: inert (digest_match = 0 and signer != sender)
    pad = 4
    inst = np.concatenate([inst, np.full(pad, I - 1, np.uint32)])
    sender = np.concatenate([sender, np.zeros(pad, np.uint16)]); signer = np.concatenate([signer, np.ones(pad, np.uint16)])
    digest_match = np.concatenate([digest_match, np.zeros(pad, np.uint8)])
    r = np.concatenate([r, np.zeros((pad, 32), np.uint8)]); s = np.concatenate([s, np.zeros((pad, 32), np.uint8)])
    dig = np.concatenate([dig, np.zeros((pad, 32), np.uint8)])
    assert inst.size == 262144
    key_of = signer.astype(np.int64) % N                              # the signature is checked against Signature.Signer's key
    qx, qy = np.ascontiguousarray(kxy[key_of, :32]), np.ascontiguousarray(kxy[key_of, 32:])
    want_ok = oracle.verify_batch(P256, r, s, qx, qy, dig)
    got_ok = eng.verify_batch(P256, r, s, qx, qy, dig)
    assert (want_ok == got_ok).all()
    # the registered-key entry point gives the same verdicts for the 16 consenters
    eng.set_keys(np.zeros(N, np.uint8), kxy.reshape(N, 2, 32))
    assert (eng.verify_registered(P256, key_of.astype(np.uint32), r, s, dig) == want_ok).all()
    cnt, reached = eng.quorum(inst, sender, signer, digest_match, got_ok, I, q - 1, self_id=np.zeros(I, np.uint16))
    want_cnt = np.zeros(I, np.int64)
    for i in range(I):
        lo, hi = i * votes_per, (i + 1) * votes_per + (pad if i == I - 1 else 0)
        if i == I - 1:
            idx = np.concatenate([np.arange(lo, (i + 1) * votes_per), np.arange(n, n + pad)])
        else:
            idx = np.arange(lo, hi)
        want_cnt[i] = ref.count_commit_votes(zip(sender[idx].tolist(), signer[idx].tolist(), digest_match[idx].tolist(), want_ok[idx].tolist()), self_id=0)
    assert cnt.tolist() == want_cnt.tolist()
    assert reached.tolist() == (want_cnt >= q - 1).astype(np.uint8).tolist()
    # at most f = 5 Byzantine votes per instance can never block a decision: Q-1 = 10 of 15 remain
    assert reached.all() and cnt.min() >= q - 1 and cnt.max() == votes_per
    # one call: signatures verified and counted with the verdicts staying on the device (sbv_verify_quorum)
    ok1, cnt1, reached1 = eng.verify_quorum(P256, r, s, qx, qy, dig, inst, sender, signer, digest_match, I, q - 1, self_id=np.zeros(I, np.uint16))
    assert (ok1 == want_ok).all() and cnt1.tolist() == want_cnt.tolist() and reached1.tolist() == reached.tolist()
    # a stricter threshold separates the instances; thresholds are monotone
    _, reached13 = eng.quorum(inst, sender, signer, digest_match, got_ok, I, 13, self_id=np.zeros(I, np.uint16))
    assert reached13.tolist() == (want_cnt >= 13).astype(np.uint8).tolist()
    assert 0 < reached13.sum() < I and (reached >= reached13).all()


def test_c5_mixed_curve_64k(eng):
    n = 65536
    tag = (np.array([corpus.DRBG(7).block(i)[0] & 1 for i in range(n)], np.uint8))
    n256, n384 = int((tag == 0).sum()), int((tag == 1).sum())
    b256 = corpus.make_batch(P256, n=n256, K=512, seed=91)
    b384 = corpus.make_batch(P384, n=n384, K=512, seed=93)
    def slot(k):
        out = np.zeros((n, 48), np.uint8)
        out[tag == 0, 16:] = b256[k]; out[tag == 1] = b384[k]
        return out
    dig = np.zeros((n, 32), np.uint8); dig[tag == 0] = b256["digest"]; dig[tag == 1] = b384["digest"]
    want = np.zeros(n, np.uint8)
    want[tag == 0] = oracle.verify_batch(P256, b256["r"], b256["s"], b256["qx"], b256["qy"], b256["digest"])
    want[tag == 1] = oracle.verify_batch(P384, b384["r"], b384["s"], b384["qx"], b384["qy"], b384["digest"])
    got = eng.verify_mixed(tag, slot("r"), slot("s"), slot("qx"), slot("qy"), dig)
    assert (want == got).all()
    assert 0.45 < tag.mean() < 0.55 and 0.9 < want.mean() < 0.96
