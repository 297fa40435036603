"""GPU parity for the round-2 paths, through the C ABI, bit-exact vs the oracle:
key grouping + on-the-fly per-key tables (sbv_verify_batch), commit-vote verification + quorum in one call
(sbv_verify_quorum), prepare collection (sbv_prepare_quorum), the rank API on a single rank, and the fault
convention (a fault is never a verdict: negative rc, ok[] untouched — SURVEY §8b, view.go:839-842)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from oracle import P256, P384, corpus
from oracle import ecdsa_ref as ref

pytestmark = pytest.mark.gpu


def _engine(**env):
    import consensus_b200 as sbv
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return sbv.Engine(n_devices=1)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.fixture(scope="module")
def eng():
    e = _engine()
    yield e
    e.close()


@pytest.mark.parametrize("thr,maxk", [(16, 8192), (1, 8192), (2, 5), (0, 8192), (100000, 8192)])
def test_grouped_path_equals_oracle_whatever_the_policy(thr, maxk):
    """Same verdicts whether every key gets a table (threshold 1: also the corrupted, off-curve keys), the table slots
    run out (5 slots), grouping is off (0), or no key is frequent enough."""
    e = _engine(SBV_GROUP_THRESHOLD=thr, SBV_GROUP_MAX_KEYS=maxk)
    try:
        for curve, n, K, seed in [(P256, 6000, 37, 31), (P384, 1500, 11, 32)]:
            b = corpus.make_batch(curve, n=n, K=K, seed=seed, corrupt_rate=5)
            want = oracle.verify_batch(curve, b["r"], b["s"], b["qx"], b["qy"], b["digest"])
            got = e.verify_batch(curve, b["r"], b["s"], b["qx"], b["qy"], b["digest"])
            bad = np.nonzero(want != got)[0]
            assert bad.size == 0, (curve, bad[:10], b["labels"][bad[:10]], want[bad[:10]], got[bad[:10]])
            assert 0 < want.sum() < n
    finally:
        e.close()


def test_all_distinct_keys_take_the_generic_path(eng):
    n = 3000
    b = corpus.make_batch(P256, n=n, K=n, seed=33, corrupt_rate=7)
    want = oracle.verify_batch(P256, b["r"], b["s"], b["qx"], b["qy"], b["digest"])
    assert np.array_equal(eng.verify_batch(P256, b["r"], b["s"], b["qx"], b["qy"], b["digest"]), want)


def test_one_key_many_signatures(eng):
    """The consensus shape: very few keys, thousands of signatures each (warp-aggregated counting, one table)."""
    b = corpus.make_batch(P256, n=40000, K=2, seed=34, corrupt_rate=9)
    want = oracle.verify_batch(P256, b["r"], b["s"], b["qx"], b["qy"], b["digest"])
    assert np.array_equal(eng.verify_batch(P256, b["r"], b["s"], b["qx"], b["qy"], b["digest"]), want)


def test_repeated_calls_reuse_scratch_sets(eng):
    """More launches than scratch sets, alternating sizes and curves (buffers grow, tables of the previous user must not leak)."""
    for i, (curve, n, K) in enumerate([(P256, 2048, 8), (P384, 600, 3), (P256, 9000, 300), (P256, 64, 1), (P384, 2000, 40), (P256, 2048, 8), (P256, 5000, 5)]):
        b = corpus.make_batch(curve, n=n, K=K, seed=40 + i, corrupt_rate=6)
        want = oracle.verify_batch(curve, b["r"], b["s"], b["qx"], b["qy"], b["digest"])
        assert np.array_equal(eng.verify_batch(curve, b["r"], b["s"], b["qx"], b["qy"], b["digest"]), want), i


def _vote_stream(I, NV, seed):
    n = I * NV
    b = corpus.make_batch(P256, n=n, K=NV + 1, seed=seed, corrupt_rate=11)
    rng = np.random.default_rng(seed)
    inst = np.repeat(np.arange(I, dtype=np.uint32), NV)
    sender = np.tile(np.arange(1, NV + 1, dtype=np.uint16), I)
    signer = sender.copy()
    dm = np.ones(n, np.uint8)
    for v in rng.choice(n, size=n // 9, replace=False):
        kind = int(rng.integers(0, 3))
        if kind == 0:
            dm[v] = 0
        elif kind == 1 and v % NV:
            sender[v] = sender[v - 1]; signer[v] = signer[v - 1]
        else:
            signer[v] = signer[v] % NV + 1 if NV > 1 else signer[v]
    return b, inst, sender, signer, dm


def test_verify_quorum_bit_exact(eng):
    """sbv_verify_quorum = verifyVote + processCommits (view.go:519-551, 827-849): verdicts stay on the device."""
    I, NV = 700, 15
    b, inst, sender, signer, dm = _vote_stream(I, NV, 51)
    sig_ok = oracle.verify_batch(P256, b["r"], b["s"], b["qx"], b["qy"], b["digest"])
    self_id = np.zeros(I, np.uint16)
    want_cnt, want_reached = ref.count_commit_votes_batch(inst, sender, signer, dm, sig_ok, I, 10, self_id)
    ok, cnt, reached = eng.verify_quorum(P256, b["r"], b["s"], b["qx"], b["qy"], b["digest"], inst, sender, signer, dm, I, 10, self_id=self_id)
    assert np.array_equal(ok, sig_ok)
    assert np.array_equal(cnt, want_cnt) and np.array_equal(reached, want_reached)
    assert 0 < reached.sum() < I
    # agrees with the two-call form (verdicts through the host)
    cnt2, reached2 = eng.quorum(inst, sender, signer, dm, sig_ok, I, 10, self_id=self_id)
    assert np.array_equal(cnt2, want_cnt) and np.array_equal(reached2, want_reached)


def test_verify_quorum_rejects_unsorted_instances(eng):
    b, inst, sender, signer, dm = _vote_stream(4, 3, 52)
    inst = inst[::-1].copy()
    import consensus_b200 as sbv
    with pytest.raises(sbv.EngineFault):
        eng.verify_quorum(P256, b["r"], b["s"], b["qx"], b["qy"], b["digest"], inst, sender, signer, dm, 4, 2)


def test_prepare_quorum_matches_processPrepares(eng):
    """view.go:441-517: Q-1 matching prepares from distinct foreign senders; the first prepare of a sender burns its slot."""
    I, NV = 500, 6
    rng = np.random.default_rng(53)
    inst = np.repeat(np.arange(I, dtype=np.uint32), NV)
    sender = rng.integers(0, 7, size=I * NV).astype(np.uint16)     # duplicates and self (0) on purpose
    dm = (rng.random(I * NV) < 0.8).astype(np.uint8)
    self_id = np.zeros(I, np.uint16)
    want_cnt, want_reached = ref.count_commit_votes_batch(inst, sender, sender, dm, np.ones(I * NV, np.uint8), I, 3, self_id)
    cnt, reached = eng.prepare_quorum(inst, sender, dm, I, 3, self_id=self_id)
    assert np.array_equal(cnt, want_cnt) and np.array_equal(reached, want_reached)
    assert 0 < reached.sum() < I


def test_rank_api_on_a_single_rank(eng):
    """sbv_verify_batch_ranked with nranks = 1: verdict bytes + the packed mask (no communicator needed)."""
    import torch
    n = 5000
    b = corpus.make_batch(P256, n=n, K=9, seed=54, corrupt_rate=4)
    want = oracle.verify_batch(P256, b["r"], b["s"], b["qx"], b["qy"], b["digest"])
    ok = np.full(n, 9, np.uint8)
    mask = np.zeros((n + 31) // 32, np.uint32)
    p = lambda a: a.ctypes.data
    f = [np.ascontiguousarray(b[k]) for k in ("r", "s", "qx", "qy", "digest")]
    eng.verify_batch_ranked_ptr(0, P256, n, *map(p, f), 32, p(ok), p(mask))
    assert np.array_equal(ok, want)
    bits = np.unpackbits(mask.view(np.uint8), bitorder="little")[:n]
    assert np.array_equal(bits, want)


# ---- a fault is never a verdict -------------------------------------------------------------------------------
def _raw(eng, name, *args):
    return getattr(eng._lib, name)(eng._h, *args)


def test_faults_return_negative_and_leave_verdicts_untouched(eng):
    n = 64
    b = corpus.make_batch(P256, n=n, K=2, seed=55, corrupt_rate=0)
    f = [np.ascontiguousarray(b[k]) for k in ("r", "s", "qx", "qy", "digest")]
    p8 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint8))
    ok = np.full(n, 0x5A, np.uint8)
    # bad curve tag
    assert _raw(eng, "sbv_verify_batch", C.c_uint8(7), C.c_size_t(n), *map(p8, f), C.c_uint8(32), p8(ok)) < 0
    # digest_len not a multiple of 4 / zero / too long
    for dl in (0, 30, 68):
        assert _raw(eng, "sbv_verify_batch", C.c_uint8(0), C.c_size_t(n), *map(p8, f), C.c_uint8(dl), p8(ok)) < 0
    # n > 2^31
    assert _raw(eng, "sbv_verify_batch", C.c_uint8(0), C.c_size_t(1 << 31), *map(p8, f), C.c_uint8(32), p8(ok)) < 0
    # null buffers
    assert _raw(eng, "sbv_verify_batch", C.c_uint8(0), C.c_size_t(n), None, p8(f[1]), p8(f[2]), p8(f[3]), p8(f[4]), C.c_uint8(32), p8(ok)) < 0
    assert _raw(eng, "sbv_verify_batch", C.c_uint8(0), C.c_size_t(n), *map(p8, f), C.c_uint8(32), None) < 0
    # decreasing offsets
    msgs = np.zeros(1024, np.uint8)
    off = np.array([0, 100, 50, 200], np.uint64)
    dig = np.full((3, 32), 0x5A, np.uint8)
    assert _raw(eng, "sbv_sha256_batch", C.c_size_t(3), p8(msgs), off.ctypes.data_as(C.POINTER(C.c_uint64)), p8(dig)) < 0
    soff = np.array([0, 70, 60, 140], np.uint32)
    assert _raw(eng, "sbv_verify_batch_der", C.c_uint8(0), C.c_size_t(3), p8(msgs), soff.ctypes.data_as(C.POINTER(C.c_uint32)), p8(msgs), p8(dig), C.c_uint8(32), p8(ok)) < 0
    assert (ok == 0x5A).all() and (dig == 0x5A).all(), "a fault must not write verdicts"
    assert b"" != eng._lib.sbv_last_error(eng._h)
    # the engine is still usable afterwards, and a real call overwrites every verdict
    want = oracle.verify_batch(P256, *f)
    assert np.array_equal(eng.verify_batch(P256, *f), want)


def test_concurrent_callers_three_lanes(eng):
    import threading
    batches = [corpus.make_batch(P256, n=3000 + 500 * i, K=4 + i, seed=60 + i, corrupt_rate=5) for i in range(5)]
    wants = [oracle.verify_batch(P256, b["r"], b["s"], b["qx"], b["qy"], b["digest"]) for b in batches]
    errs = []

    def work(i):
        try:
            for _ in range(3):
                b = batches[i]
                got = eng.verify_batch(P256, b["r"], b["s"], b["qx"], b["qy"], b["digest"])
                if not np.array_equal(got, wants[i]):
                    errs.append(i)
        except Exception as ex:
            errs.append(repr(ex))
    ths = [threading.Thread(target=work, args=(i,)) for i in range(5)]
    for t in ths: t.start()
    for t in ths: t.join()
    assert not errs, errs


# ---- single-process multi-device engine (sbv_create with 2 devices): needs a 2-GPU box ----------------------------
def _two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")


def test_multi_device_verify_quorum_shards_by_instance():
    """sbv_verify_quorum on a 2-device engine: votes sharded by instance, verdict + reached masks in one NCCL all-gather."""
    _two_gpus()
    import consensus_b200 as sbv
    I, NV = 901, 15      # odd instance count: uneven shards
    b, inst, sender, signer, dm = _vote_stream(I, NV, 71)
    sig_ok = oracle.verify_batch(P256, b["r"], b["s"], b["qx"], b["qy"], b["digest"])
    self_id = np.zeros(I, np.uint16)
    want_cnt, want_reached = ref.count_commit_votes_batch(inst, sender, signer, dm, sig_ok, I, 10, self_id)
    with sbv.Engine(n_devices=2) as e2:
        for _ in range(3):
            ok, cnt, reached = e2.verify_quorum(P256, b["r"], b["s"], b["qx"], b["qy"], b["digest"], inst, sender, signer, dm, I, 10, self_id=self_id)
            assert np.array_equal(ok, sig_ok)
            assert np.array_equal(cnt, want_cnt) and np.array_equal(reached, want_reached)


def test_multi_device_concurrent_callers():
    """Two host threads on a 2-device engine: per-lane gather buffers, collectives issued under one order."""
    _two_gpus()
    import threading
    import consensus_b200 as sbv
    batches = [corpus.make_batch(P256, n=4001 + 777 * i, K=5 + i, seed=80 + i, corrupt_rate=5) for i in range(3)]
    wants = [oracle.verify_batch(P256, b["r"], b["s"], b["qx"], b["qy"], b["digest"]) for b in batches]
    errs = []
    with sbv.Engine(n_devices=2) as e2:
        def work(i):
            try:
                for _ in range(4):
                    b = batches[i]
                    if not np.array_equal(e2.verify_batch(P256, b["r"], b["s"], b["qx"], b["qy"], b["digest"]), wants[i]):
                        errs.append(i)
            except Exception as ex:
                errs.append(repr(ex))
        ths = [threading.Thread(target=work, args=(i,)) for i in range(3)]
        for t in ths: t.start()
        for t in ths: t.join()
    assert not errs, errs


def _signed_requests(curve, n, K, seed, lo=40, hi=700):
    msgs, off = corpus.make_requests(n, seed=seed, fixed_len=None, lo=lo, hi=hi)
    d, kxy = corpus.make_keys(curve, K, seed=seed + 1)
    L = 32 if curve == P256 else 48
    key_idx = (np.arange(n) * 7 % K).astype(np.uint32)
    dig = oracle.sha256_batch(msgs, off)
    r, s = oracle.sign_batch(curve, d, key_idx, dig, corpus._blocks(seed + 2, n, L, b"k"))
    qx, qy = np.ascontiguousarray(kxy[key_idx, :L]), np.ascontiguousarray(kxy[key_idx, L:])
    msgs = msgs.copy()
    for i in range(0, n, 5):
        msgs[int(off[i]) + (i % int(off[i + 1] - off[i]))] ^= 1     # payload bit
    s[::7, 9] ^= 2                                                   # signature bit
    qx[::11] = qx[3]                                                 # another signer's key
    return msgs, off, r, s, qx, qy


@pytest.mark.parametrize("thr", [16, 0, 1])
def test_chunked_upload_equals_oracle(thr):
    """A shard of >= 2 x SBV_CHUNK_ITEMS items arrives and is verified chunk by chunk (shared grouping and key tables,
    chunk-local routing): same verdicts and digests as the oracle, for a chunk size that does not divide the batch, with
    grouping on, off, and with a table for every key — from pageable and from pinned caller memory."""
    import torch
    e = _engine(SBV_CHUNK_ITEMS=1000, SBV_GROUP_THRESHOLD=thr)
    try:
        for curve, n, K in [(P256, 7013, 41), (P384, 2300, 9), (P256, 2000, 2000)]:
            msgs, off, r, s, qx, qy = _signed_requests(curve, n, K, seed=900 + n)
            want_dig = oracle.sha256_batch(msgs, off)
            want = oracle.verify_batch(curve, r, s, qx, qy, want_dig)
            assert 0 < want.sum() < n
            got, got_dig = e.hash_verify_batch(curve, msgs, off, r, s, qx, qy, want_digest=True)
            assert np.array_equal(got_dig, want_dig)
            assert np.array_equal(got, want), (curve, n, np.nonzero(got != want)[0][:10])
            assert np.array_equal(e.verify_batch(curve, r, s, qx, qy, want_dig), want)
            # pinned caller memory: the chunks are copied straight from the caller's buffers on the upload stream
            pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
            assert np.array_equal(e.verify_batch(curve, pin(r), pin(s), pin(qx), pin(qy), pin(want_dig)), want)
            assert np.array_equal(e.hash_verify_batch(curve, pin(msgs), pin(off), pin(r), pin(s), pin(qx), pin(qy)), want)
        # back to a small (unchunked) call on the same engine and lanes
        b = corpus.make_batch(P256, n=500, K=4, seed=77, corrupt_rate=5)
        assert np.array_equal(e.verify_batch(P256, b["r"], b["s"], b["qx"], b["qy"], b["digest"]),
                              oracle.verify_batch(P256, b["r"], b["s"], b["qx"], b["qy"], b["digest"]))
    finally:
        e.close()


def test_chunked_upload_many_chunks_and_concurrent_callers():
    """More chunks than SBV_MAX_CHUNKS would allow at the nominal size (clamped), from several caller threads at once."""
    import threading
    e = _engine(SBV_CHUNK_ITEMS=64)
    try:
        jobs = []
        for t in range(3):
            msgs, off, r, s, qx, qy = _signed_requests(P256, 5000 + 300 * t, 13 + t, seed=950 + t)
            jobs.append((msgs, off, r, s, qx, qy, oracle.verify_batch(P256, r, s, qx, qy, oracle.sha256_batch(msgs, off))))
        out = [None] * len(jobs)

        def run(i):
            m, o, r, s, qx, qy, _ = jobs[i]
            for _ in range(3):
                out[i] = e.hash_verify_batch(P256, m, o, r, s, qx, qy)
        th = [threading.Thread(target=run, args=(i,)) for i in range(len(jobs))]
        [t.start() for t in th]
        [t.join() for t in th]
        for i, j in enumerate(jobs):
            assert np.array_equal(out[i], j[6]), i
    finally:
        e.close()
