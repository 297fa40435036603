"""oracle/corpus.py — seeded synthetic corpora.  TEST / BENCH INPUT GENERATOR.

All bytes derive from a counter-mode SHA-256 DRBG, so every process (tests, bench ranks, the CPU
baseline) regenerates identical inputs from the seed.  Expected verdicts are never taken from the
corruption label — always from the oracle.
"""
from __future__ import annotations

import hashlib

import numpy as np

from . import FIELD_BYTES, P256, P384, pubkey, sign_batch
from .ecdsa_ref import CURVES, DRBG

N_CLASSES = 12
CLASS_NAMES = ["flip_r", "flip_s", "flip_e", "flip_qx", "swap_key", "r_zero", "s_zero", "r_eq_n",
               "s_eq_n_plus_1", "r_plus_n", "q_zero", "high_s"]


def _blocks(seed: int, n: int, width: int, tag: bytes = b"") -> np.ndarray:
    """n rows of `width` DRBG bytes."""
    pre = str(seed).encode() + b":" + tag
    out = np.empty((n, width), np.uint8)
    reps = (width + 31) // 32
    for i in range(n):
        b = b"".join(hashlib.sha256(pre + str(i).encode() + b"/" + str(j).encode()).digest() for j in range(reps))
        out[i] = np.frombuffer(b[:width], np.uint8)
    return out


def make_keys(curve: int, K: int, seed: int = 1):
    c = CURVES[curve]
    L = FIELD_BYTES[curve]
    raw = _blocks(seed, K, L + 8, b"key")
    d = np.zeros((K, L), np.uint8)
    keys = np.zeros((K, 2 * L), np.uint8)
    for k in range(K):
        dv = int.from_bytes(raw[k].tobytes(), "big") % (c.n - 1) + 1
        db = dv.to_bytes(L, "big")
        d[k] = np.frombuffer(db, np.uint8)
        qx, qy = pubkey(curve, db)
        keys[k, :L] = np.frombuffer(qx, np.uint8)
        keys[k, L:] = np.frombuffer(qy, np.uint8)
    return d, keys


def make_digests(n: int, seed: int = 2) -> np.ndarray:
    pre = b"msg" + str(seed).encode() + b":"
    out = np.empty((n, 32), np.uint8)
    for i in range(n):
        out[i] = np.frombuffer(hashlib.sha256(pre + str(i).encode()).digest(), np.uint8)
    return out


def corrupt(curve: int, batch: dict, seed: int = 4, rate: int = 16) -> np.ndarray:
    """In-place corruption of ≈1/rate of the items over the 12 classes.  Returns the label array
    (-1 = untouched)."""
    c = CURVES[curve]
    L = FIELD_BYTES[curve]
    n = batch["r"].shape[0]
    g = DRBG(seed)
    labels = np.full(n, -1, np.int16)
    K = batch["keys"].shape[0]
    for i in range(n):
        blk = g.block(i)
        if int.from_bytes(blk[:4], "big") % rate:
            continue
        cls = int.from_bytes(blk[4:8], "big") % N_CLASSES
        bit = int.from_bytes(blk[8:12], "big")
        labels[i] = cls
        flip = lambda a: a.__setitem__((i, (bit >> 3) % a.shape[1]), a[i, (bit >> 3) % a.shape[1]] ^ (1 << (bit & 7)))
        if cls == 0:
            flip(batch["r"])
        elif cls == 1:
            flip(batch["s"])
        elif cls == 2:
            flip(batch["digest"])
        elif cls == 3:
            flip(batch["qx"])
        elif cls == 4:
            k2 = (int(batch["key_idx"][i]) + 1 + bit % max(K - 1, 1)) % K
            batch["qx"][i] = batch["keys"][k2, :L]
            batch["qy"][i] = batch["keys"][k2, L:]
        elif cls == 5:
            batch["r"][i] = 0
        elif cls == 6:
            batch["s"][i] = 0
        elif cls == 7:
            batch["r"][i] = np.frombuffer(c.n.to_bytes(L, "big"), np.uint8)
        elif cls == 8:
            batch["s"][i] = np.frombuffer((c.n + 1).to_bytes(L, "big"), np.uint8)
        elif cls == 9:
            rv = int.from_bytes(batch["r"][i].tobytes(), "big") + c.n
            if rv < (1 << (8 * L)):
                batch["r"][i] = np.frombuffer(rv.to_bytes(L, "big"), np.uint8)
            else:
                flip(batch["r"])
        elif cls == 10:
            batch["qx"][i] = 0
            batch["qy"][i] = 0
        elif cls == 11:
            sv = c.n - int.from_bytes(batch["s"][i].tobytes(), "big")
            batch["s"][i] = np.frombuffer(sv.to_bytes(L, "big"), np.uint8)
    return labels


def make_batch(curve: int = P256, n: int = 65536, K: int = 1024, seed: int = 1, corrupt_rate: int = 16,
               keys=None) -> dict:
    """C2-style batch: n signatures over K keys, ≈1/16 corrupted.  SoA big-endian arrays."""
    L = FIELD_BYTES[curve]
    d, kxy = keys if keys is not None else make_keys(curve, K, seed)
    K = kxy.shape[0]
    digest = make_digests(n, seed + 1)
    key_idx = (np.arange(n, dtype=np.uint32) % K).astype(np.uint32)
    nonces = _blocks(seed + 2, n, L, b"k")
    r, s = sign_batch(curve, d, key_idx, digest, nonces)
    batch = {
        "curve": curve, "n": n, "r": r, "s": s,
        "qx": np.ascontiguousarray(kxy[key_idx, :L]), "qy": np.ascontiguousarray(kxy[key_idx, L:]),
        "digest": digest, "key_idx": key_idx, "keys": kxy, "priv": d,
    }
    batch["labels"] = corrupt(curve, batch, seed + 3, corrupt_rate) if corrupt_rate else np.full(n, -1, np.int16)
    return batch


def make_requests(n: int, seed: int = 5, fixed_len: int | None = 256, lo: int = 64, hi: int = 10240):
    """C3 request bytes: concatenated payload + uint64 offsets[n+1]."""
    g = DRBG(seed)
    if fixed_len is not None:
        lens = np.full(n, fixed_len, np.int64)
    else:
        # log-uniform in [lo, hi]
        u = np.array([g.u32(i) for i in range(n)], np.float64) / 2.0**32
        lens = np.floor(np.exp(np.log(lo) + u * (np.log(hi + 1) - np.log(lo)))).astype(np.int64)
        lens = np.clip(lens, lo, hi)
    off = np.zeros(n + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    total = int(off[-1])
    # bulk bytes: a seeded PCG stream is used for volume (the DRBG only seeds it) — 1M × 256 B of
    # SHA-256 counter output would dominate test time for no extra coverage.
    rng = np.random.Generator(np.random.PCG64(int.from_bytes(g.block(0)[:8], "big")))
    msgs = rng.integers(0, 256, size=total, dtype=np.uint8)
    return msgs, off
