"""CPU simulation of the DEVICE code (tools/hostsim: mp.cuh / curve.cuh / kernels.cuh / keygroup.cuh / sha256.cuh / quorum.cuh
compiled with g++, PTX carry-flag primitives emulated; thread-per-item kernels one simulated thread at a time,
warp-cooperative kernels in lockstep with one OS thread per lane) against Python big integers, hashlib and the oracle.

This is how limb-level arithmetic and the kernels' control flow are checked without a GPU; the `-m gpu` tests run the
same checks on the real thing.  The simulation is test infrastructure — libsbv.so has no CPU path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from oracle import corpus
from oracle import ecdsa_ref as ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HS_DIR = os.path.join(ROOT, "tools", "hostsim")


@pytest.fixture(scope="module")
def hs():
    subprocess.check_call(["make", "-s", "-C", HS_DIR, "libhostsim.so"])
    return C.CDLL(os.path.join(HS_DIR, "libhostsim.so"))


def _limbs(vals, N):
    out = np.zeros((len(vals), 2 * N), np.uint32)
    for i, pair in enumerate(vals):
        for h, v in enumerate(pair):
            for k in range(N):
                out[i, h * N + k] = (v >> (32 * k)) & 0xFFFFFFFF
    return out


def _ints(arr, N):
    return [tuple(sum(int(row[h * N + k]) << (32 * k) for k in range(N)) for h in range(2)) for row in arr]


def _run(hs, curve, op, a, b):
    N = 8 if curve == 0 else 12
    A, B = _limbs(a, N), _limbs(b, N)
    out = np.zeros_like(A)
    p32 = lambda x: x.ctypes.data_as(C.POINTER(C.c_uint32))
    assert hs.hs_debug_op(C.c_int(curve), C.c_int(op), C.c_size_t(len(a)), p32(A), p32(B), p32(out)) == 0
    return _ints(out, N)


def _edge_values(m, rng, count):
    F = (1 << 32) - 1
    vals = [0, 1, 2, m - 1, m - 2, (m - 1) // 2, F, 1 << 32, (1 << 64) - 1, 1 << 96, (1 << 224) % m, m >> 1,
            0xFFFFFFFF00000000FFFFFFFF00000000FFFFFFFF00000000FFFFFFFF00000000 % m, ((1 << 96) - 1), (m - (1 << 96)) % m,
            (m - (1 << 192)) % m, ((1 << 256) - 1) % m, ((1 << 255) + 12345) % m]
    vals += [int.from_bytes(rng.bytes(48), "big") % m for _ in range(count)]
    # values with long runs of ones / zeros in single limbs: carry-chain stress
    for _ in range(count // 4):
        v = 0
        for k in range(12):
            v |= int(rng.choice([0, F, 1, F - 1, 0x80000000, int(rng.integers(0, F))])) << (32 * k)
        vals.append(v % m)
    return vals


@pytest.mark.parametrize("curve", [0, 1])
def test_field_ops(hs, curve):
    c = ref.CURVES[curve]
    N = c.size // 4
    R = 1 << (32 * N)
    rng = np.random.default_rng(curve + 11)
    for m, mulop in [(c.p, 0), (c.n, 3)]:
        xs = _edge_values(m, rng, 1500)
        ys = list(reversed(_edge_values(m, rng, 1500)))
        a = [(x, 0) for x in xs]; b = [(y, 0) for y in ys]
        Rinv = pow(R, -1, m)
        assert [g[0] for g in _run(hs, curve, mulop, a, b)] == [x * y * Rinv % m for x, y in zip(xs, ys)]
        if m == c.p:
            assert [g[0] for g in _run(hs, curve, 9, a, b)] == [x * x * Rinv % m for x in xs]
            assert [g[0] for g in _run(hs, curve, 1, a, b)] == [(x + y) % m for x, y in zip(xs, ys)]
            assert [g[0] for g in _run(hs, curve, 2, a, b)] == [(x - y) % m for x, y in zip(xs, ys)]
    xs = [v for v in _edge_values(c.p, rng, 10) if v]
    got = _run(hs, curve, 4, [(x * R % c.p, 0) for x in xs], [(0, 0)] * len(xs))
    assert [g[0] for g in got] == [pow(x, -1, c.p) * R % c.p for x in xs]
    xs = [v for v in _edge_values(c.p, rng, 300) if v] + [pow(2, k, c.p) for k in (1, 31, 32, 33, 64, 96, 128, 224, 255, 256, 300)]
    got = _run(hs, curve, 10, [(x * R % c.p, 0) for x in xs], [(0, 0)] * len(xs))   # binary-GCD field inverse
    assert [g[0] for g in got] == [pow(x, -1, c.p) * R % c.p for x in xs]
    xs = [v for v in _edge_values(c.n, rng, 200) if v]
    xs += [pow(2, k, c.n) for k in (1, 31, 32, 33, 63, 64, 65, 96, 128, 255, 256, 300, 383)]
    got = _run(hs, curve, 8, [(x * R % c.n, 0) for x in xs], [(0, 0)] * len(xs))
    assert [g[0] for g in got] == [pow(x, -1, c.n) * R % c.n for x in xs]


@pytest.mark.parametrize("curve", [0, 1])
def test_group_law(hs, curve):
    c = ref.CURVES[curve]
    G = (c.gx, c.gy)
    ks = [1, 2, 3, 4, 5, 7, 8, 255, c.n - 1, c.n - 2, 2**100 + 3]
    pts = [ref.scalar_mult(c, k, G) for k in ks]
    assert _run(hs, curve, 5, pts, pts) == [ref._add(c, P, P) for P in pts]
    pairs = [(P, Q) for P in pts for Q in pts]
    want = [ref._add(c, P, Q) or (0, 0) for P, Q in pairs]
    assert _run(hs, curve, 7, [p for p, _ in pairs], [q for _, q in pairs]) == want
    want = [ref._add(c, ref._add(c, P, P), Q) or (0, 0) for P, Q in pairs]
    assert _run(hs, curve, 6, [p for p, _ in pairs], [q for _, q in pairs]) == want


def _p8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _verify(hs, curve, b, grouped=None):
    n = b["r"].shape[0]
    ok = np.full(n, 7, np.uint8)
    f = [np.ascontiguousarray(b[k]) for k in ("r", "s", "qx", "qy", "digest")]
    dlen = f[4].size // n
    if grouped is None:
        assert hs.hs_verify(C.c_int(curve), C.c_size_t(n), *map(_p8, f), C.c_uint32(dlen), _p8(ok)) == 0
        return ok
    thr, maxk = grouped
    stats = np.zeros(3, np.uint32)
    assert hs.hs_verify_grouped(C.c_int(curve), C.c_size_t(n), *map(_p8, f), C.c_uint32(dlen), C.c_uint32(thr), C.c_uint32(maxk), _p8(ok),
                                stats.ctypes.data_as(C.POINTER(C.c_uint32))) == 0
    return ok, stats


def test_verify_generic_kernel_p256(hs):
    """k_prep + k_verify_coz (the kernels of the product, simulated) on a corrupted corpus: every class of the
    corpus, bit-exact with the oracle."""
    b = corpus.make_batch(oracle.P256, n=192, K=8, seed=21, corrupt_rate=3)
    want = oracle.verify_batch(oracle.P256, b["r"], b["s"], b["qx"], b["qy"], b["digest"])
    assert 0 < int(want.sum()) < want.size
    assert np.array_equal(_verify(hs, 0, b), want)


def test_verify_generic_kernel_p384(hs):
    b = corpus.make_batch(oracle.P384, n=48, K=4, seed=22, corrupt_rate=3)
    want = oracle.verify_batch(oracle.P384, b["r"], b["s"], b["qx"], b["qy"], b["digest"])
    assert np.array_equal(_verify(hs, 1, b), want)


@pytest.mark.parametrize("thr,maxk", [(4, 64), (1, 64), (4, 3), (1000, 64)])
def test_verify_grouped_p256(hs, thr, maxk):
    """Key grouping + on-the-fly per-key tables + fixed-base kernel, with the generic kernel for the rest: same
    verdicts as the oracle whatever the threshold / table capacity (all keys grouped, capacity exhausted, nothing
    grouped)."""
    b = corpus.make_batch(oracle.P256, n=256, K=8, seed=23, corrupt_rate=3)
    want = oracle.verify_batch(oracle.P256, b["r"], b["s"], b["qx"], b["qy"], b["digest"])
    got, stats = _verify(hs, 0, b, grouped=(thr, maxk))
    assert np.array_equal(got, want)
    assert int(stats[1]) + int(stats[2]) == 256
    if thr == 1000:
        assert int(stats[1]) == 0
    if thr == 1 and maxk == 64:
        assert int(stats[2]) == 0          # every key (also the corrupted, off-curve ones) got a slot; invalid ones reject by flag
    if maxk == 3:
        assert int(stats[0]) >= 3 and int(stats[2]) > 0


def test_verify_grouped_p384(hs):
    b = corpus.make_batch(oracle.P384, n=64, K=3, seed=24, corrupt_rate=3)
    want = oracle.verify_batch(oracle.P384, b["r"], b["s"], b["qx"], b["qy"], b["digest"])
    got, stats = _verify(hs, 1, b, grouped=(4, 16))
    assert np.array_equal(got, want)
    assert int(stats[1]) > 0


@pytest.mark.parametrize("curve,n,K,thr,chunk", [(0, 250, 8, 4, 96), (0, 250, 8, 3, 37), (1, 60, 3, 4, 25)])
def test_verify_chunked_second_half(hs, curve, n, K, thr, chunk):
    """The second half of the pipeline run chunk by chunk (what a chunked host-buffer call enqueues, csrc/pipeline.cu:
    sbv_launch_verify_chunk): shared grouping and key tables, per-chunk routing with chunk-local indices, every per-item
    array addressed as the contiguous slice of the chunk — same verdicts and the same split between the two paths as
    the one-piece launch, for chunk sizes that do not divide the batch, with and without the split u1*G kernel."""
    cv = oracle.P256 if curve == 0 else oracle.P384
    b = corpus.make_batch(cv, n=n, K=K, seed=25 + chunk, corrupt_rate=3)
    want = oracle.verify_batch(cv, b["r"], b["s"], b["qx"], b["qy"], b["digest"])
    whole, st_whole = _verify(hs, curve, b, grouped=(thr, 64))
    ok = np.full(n, 7, np.uint8)
    f = [np.ascontiguousarray(b[k]) for k in ("r", "s", "qx", "qy", "digest")]
    stats = np.zeros(3, np.uint32)
    assert hs.hs_verify_chunked(C.c_int(curve), C.c_size_t(n), *map(_p8, f), C.c_uint32(f[4].size // n), C.c_uint32(thr), C.c_uint32(64),
                                C.c_uint32(chunk), _p8(ok), stats.ctypes.data_as(C.POINTER(C.c_uint32))) == 0
    assert np.array_equal(ok, want)
    assert np.array_equal(whole, want)
    assert list(stats) == list(st_whole) and int(stats[1]) > 0


def _crafted(curve, cases):
    """(u1, u2, k) -> a signature on Q = k*G whose verification computes exactly u1*G + u2*Q (s = r/u2, e = u1*s):
    places exceptional points (doubling, P + (-P), infinity in the middle or at the end) inside the scalar multiplication."""
    c = ref.CURVES[curve]
    L = c.size
    rows = []
    for u1, u2, k in cases:
        Q = ref.scalar_mult(c, k % c.n, (c.gx, c.gy))
        R = ref._add(c, ref.scalar_mult(c, u1 % c.n, (c.gx, c.gy)), ref.scalar_mult(c, u2 % c.n, Q))
        if u2 % c.n == 0:
            continue
        r = 1 if R is None else R[0] % c.n
        if r == 0:
            continue
        s = r * pow(u2, -1, c.n) % c.n
        rows.append((r, s, Q[0], Q[1], u1 * s % c.n))
    f = lambda j: np.stack([np.frombuffer(int(row[j]).to_bytes(L, "big"), np.uint8) for row in rows])
    return {"r": f(0), "s": f(1), "qx": f(2), "qy": f(3), "digest": f(4)}


@pytest.mark.parametrize("curve,thr", [(0, 1), (0, 2), (1, 2)])
def test_exceptional_points_on_the_fixed_base_path(hs, curve, thr):
    """Every key gets a table (threshold 1 / 2: with the u1*G half inside the fixed-base kernel and in k_gpart) and the
    scalars are chosen so that the running sum meets the next table entry (doubling inside a mixed addition), its
    negative (infinity in the middle), or ends at infinity (must reject)."""
    c = ref.CURVES[curve]
    n = c.n
    ks = [1, 2, 3, n - 1, 5, 2**8 + 1] if curve == 0 else [1, 3, n - 2]
    cases = []
    for k in ks:
        for u1, u2 in [(1, 1), (k, 1), (n - k, 1), (k, n - 1), (2, n - 1), (1, 2), (7, 3), (2**255, 2**255), (n - 1, n - 1), (k * 5 % n, 5),
                       (n - (k * 5 % n), 5), (k * 16 % n, 16), (n - (k * 16 % n), 16), (k * 33 % n, 33), (2**64, 2**64), (16, 1), (1, 16),
                       (0, 1), (0, 77), ((k << 5) % n, 32), (n - ((k << 5) % n), 32)]:
            cases.append((u1, u2, k))
    b = _crafted(curve, cases)
    want = oracle.verify_batch(curve, b["r"], b["s"], b["qx"], b["qy"], b["digest"])
    assert 0 < int(want.sum()) < want.size
    got, stats = _verify(hs, curve, b, grouped=(thr, 64))
    assert int(stats[2]) == 0                                   # nothing on the generic path
    assert np.array_equal(got, want), np.nonzero(got != want)[0][:10]


def _tables(hs, curve, w8, kxy, four):
    L = 32 if curve == 0 else 48
    n = kxy.shape[0]
    qx, qy = np.ascontiguousarray(kxy[:, :L]), np.ascontiguousarray(kxy[:, L:])
    hs.hs_ktab_words.restype = C.c_size_t
    words = hs.hs_ktab_words(C.c_int(curve), C.c_int(w8), C.c_size_t(n))
    kt, fl = np.zeros(words, np.uint32), np.zeros(n, np.uint8)
    assert hs.hs_tables(C.c_int(curve), C.c_int(w8), C.c_size_t(n), _p8(qx), _p8(qy), C.c_int(four), kt.ctypes.data_as(C.POINTER(C.c_uint32)), _p8(fl)) == 0
    return kt, fl


@pytest.mark.parametrize("curve,w8,nkeys", [(0, 0, 6), (0, 1, 3), (1, 0, 3)])
def test_four_lane_doubling_chain_in_lockstep(hs, curve, w8, nkeys):
    """k_kt_bases4 — four lanes per key, the independent multiplications of a doubling on different lanes, quad shuffles —
    simulated with one OS thread per lane meeting at every shuffle: the tables it leads to are bit-identical to those of the
    one-thread-per-key kernel, the validity flags too (an off-curve key and a coordinate >= p among the keys), and every
    entry of a table is e * 2^(W*w) * Q in affine Montgomery form."""
    cv = oracle.P256 if curve == 0 else oracle.P384
    c = ref.CURVES[cv]
    L, N, W = c.size, c.size // 4, 8 if w8 else 5
    d, kxy = corpus.make_keys(cv, nkeys, seed=3 + curve)
    kxy = kxy.copy()
    kxy[1, L + 8] ^= 1                                            # off the curve
    if nkeys > 3:
        kxy[4, :L] = np.frombuffer(int(c.p + 2).to_bytes(L, "big"), np.uint8)   # x >= p
    a, fa = _tables(hs, curve, w8, kxy, four=0)
    b, fb = _tables(hs, curve, w8, kxy, four=1)
    assert fa.tolist() == fb.tolist() and fa[0] == 1 and fa[1] == 0 and (nkeys <= 3 or fa[4] == 0)
    assert np.array_equal(a, b)
    # key 0 against Python integers
    nwin = (8 * L + 1 + W - 1) // W
    ent = 1 << (W - 1)
    tab = b[: nwin * ent * 2 * N].reshape(nwin, ent, 2, N)
    Q = (int.from_bytes(kxy[0, :L].tobytes(), "big"), int.from_bytes(kxy[0, L:].tobytes(), "big"))
    Rm = 1 << (8 * L)
    val = lambda w: sum(int(x) << (32 * i) for i, x in enumerate(w))
    for win in sorted({0, 1, nwin // 2, nwin - 1}):
        for e in (1, 2, 3, ent - 1, ent):
            if win == nwin - 1 and (e << (W * win)) >= c.n:
                continue
            P = ref.scalar_mult(c, (e << (W * win)) % c.n, Q)
            assert val(tab[win, e - 1, 0]) == P[0] * Rm % c.p and val(tab[win, e - 1, 1]) == P[1] * Rm % c.p, (win, e)


@pytest.mark.parametrize("curve,n", [(0, 40), (1, 12)])
def test_registered_keys_thread_and_warp_kernels(hs, curve, n):
    """sbv_set_keys / sbv_verify_registered on the CPU: 8-bit-window tables, keys taken by slot, the thread-per-signature
    kernel and the ONE-SIGNATURE-PER-WARP kernel (32 OS threads per signature, shuffle-tree reduction in lockstep):
    both equal the oracle, including an invalid registered key and an out-of-range slot."""
    cv = oracle.P256 if curve == 0 else oracle.P384
    c = ref.CURVES[cv]
    L, K = c.size, 4
    b = corpus.make_batch(cv, n=n, K=K - 1, seed=60 + curve, corrupt_rate=4)
    d, kxy = corpus.make_keys(cv, K, seed=70 + curve)
    key_idx = (np.arange(n) % K).astype(np.uint32)
    r, s = oracle.sign_batch(cv, d, key_idx, b["digest"], corpus._blocks(71, n, L, b"k"))
    s[::5, L // 2] ^= 4                                          # bad signatures
    kxy = kxy.copy()
    kxy[3, L + 3] ^= 8                                           # registered key 3 is not on the curve
    slot = key_idx.copy()
    slot[7] = 99                                                 # no such slot
    qx, qy = np.ascontiguousarray(kxy[key_idx, :L]), np.ascontiguousarray(kxy[key_idx, L:])
    want = oracle.verify_batch(cv, r, s, qx, qy, b["digest"])
    want[7] = 0
    assert want[key_idx == 3].sum() == 0 and 0 < want.sum() < n
    kx, ky = np.ascontiguousarray(kxy[:, :L]), np.ascontiguousarray(kxy[:, L:])
    dig = np.ascontiguousarray(b["digest"])
    for warp in (0, 1):
        ok = np.full(n, 7, np.uint8)
        assert hs.hs_verify_registered(C.c_int(curve), C.c_size_t(n), C.c_size_t(K), _p8(kx), _p8(ky), slot.ctypes.data_as(C.POINTER(C.c_uint32)),
                                       _p8(r), _p8(s), _p8(dig), C.c_uint32(dig.size // n), C.c_int(warp), _p8(ok)) == 0
        assert np.array_equal(ok, want), (warp, np.nonzero(ok != want)[0])


def test_sha256_kernel_on_a_ragged_batch(hs):
    """k_sha256 (aligned 32-bit loads re-aligned with PRMT, padding built in registers) against hashlib: every length
    0..200 plus block boundaries, at arbitrary byte offsets, in input order and in a permuted processing order."""
    import hashlib
    rng = np.random.default_rng(5)
    lens = list(range(0, 200)) + [247, 248, 255, 256, 257, 1000, 4095, 4096] + rng.integers(0, 1500, 60).tolist()
    off = np.zeros(len(lens) + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    lead = 3                                                    # the batch does not start on a word boundary
    off += np.uint64(lead)
    buf = rng.integers(0, 256, int(off[-1]) + 16, dtype=np.uint8)
    n = len(lens)
    want = np.stack([np.frombuffer(hashlib.sha256(buf[int(off[i]):int(off[i + 1])].tobytes()).digest(), np.uint8) for i in range(n)])
    for perm in (None, rng.permutation(n).astype(np.uint32)):
        out = np.zeros((n, 32), np.uint8)
        assert hs.hs_sha256(C.c_size_t(n), _p8(buf), off.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_uint64(0),
                            None if perm is None else perm.ctypes.data_as(C.POINTER(C.c_uint32)), _p8(out)) == 0
        assert np.array_equal(out, want)


def test_quorum_and_pack_kernels(hs):
    """k_quorum_count / k_quorum_reached against the oracle's restatement of processCommits (view.go:519-551, util.go:130-143)
    on a Byzantine vote stream — whole, and as the two shards of a two-device engine (instances split, instance base
    subtracted) — and k_pack_bits (a warp ballot per 32 verdicts, in lockstep) against the host-side packing."""
    from consensus_b200 import sharding
    rng = np.random.default_rng(11)
    I, NV = 37, 7
    inst = np.repeat(np.arange(I, dtype=np.uint32), NV)
    nv = inst.size
    sender = (np.tile(np.arange(NV), I) + 1).astype(np.uint16)
    dup = rng.random(nv) < 0.15
    sender[dup] = np.roll(sender, 1)[dup]                       # duplicate senders: the later vote must not count
    signer = sender.copy()
    wrong = rng.random(nv) < 0.1
    signer[wrong] = (signer[wrong] % NV) + 1                     # signer != sender: never registered (and burns nothing)
    dm = (rng.random(nv) > 0.1).astype(np.uint8)
    ok = (rng.random(nv) > 0.2).astype(np.uint8)
    self_id = (rng.integers(0, NV, I) + 1).astype(np.uint16)     # the counting node's own id per instance
    thr = 4
    want_cnt, want_rch = ref.count_commit_votes_batch(inst, sender, signer, dm, ok, I, thr, self_id)
    p16, p32 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint16)), lambda a: a.ctypes.data_as(C.POINTER(C.c_uint32))

    def run(vlo, vhi, ilo, ihi, with_ok=True):
        cnt, rch = np.zeros(ihi - ilo, np.uint32), np.zeros(ihi - ilo, np.uint8)
        a = [np.ascontiguousarray(x[vlo:vhi]) for x in (inst, sender, signer, dm, ok)]
        sid = np.ascontiguousarray(self_id[ilo:ihi])
        assert hs.hs_quorum(C.c_size_t(vhi - vlo), p32(a[0]), p16(a[1]), p16(a[2]), _p8(a[3]), _p8(a[4]) if with_ok else None, p16(sid),
                            C.c_uint32(ilo), C.c_size_t(ihi - ilo), C.c_uint32(thr), p32(cnt), _p8(rch)) == 0
        return cnt, rch
    cnt, rch = run(0, nv, 0, I)
    assert np.array_equal(cnt, want_cnt) and np.array_equal(rch, want_rch) and 0 < rch.sum() < I
    parts = []
    for g in range(2):                                           # sharded by instance, as sbv_verify_quorum does on two devices
        ilo, ihi = sharding.shard_range(I, g, 2)
        vlo, vhi = int(np.searchsorted(inst, ilo)), int(np.searchsorted(inst, ihi))
        parts.append(run(vlo, vhi, ilo, ihi))
    assert np.array_equal(np.concatenate([p[0] for p in parts]), want_cnt)
    # prepares carry no signature (ok == NULL): every registered vote with a matching digest counts
    cnt_p, _ = run(0, nv, 0, I, with_ok=False)
    want_p, _ = ref.count_commit_votes_batch(inst, sender, signer, dm, np.ones(nv, np.uint8), I, thr, self_id)
    assert np.array_equal(cnt_p, want_p)
    # verdict bytes -> bitmask
    for n in (1, 31, 32, 33, 1000):
        v = (rng.random(n) > 0.4).astype(np.uint8)
        words = np.zeros((n + 31) // 32, np.uint32)
        assert hs.hs_pack_bits(C.c_size_t(n), _p8(v), p32(words)) == 0
        assert np.array_equal(words, sharding.pack_bits(v, words.size)), n
