"""oracle — CPU oracle for the sbv hot path.  TEST INFRASTRUCTURE ONLY (never imported by consensus_b200).

`oracle.lib` wraps liboracle.so (oracle/oracle.c, OpenSSL-backed restatement, fast, multi-threaded);
`oracle.ecdsa_ref` is the independent pure-Python big-int restatement for small cases.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import ecdsa_ref  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

P256, P384 = 0, 1
FIELD_BYTES = {P256: 32, P384: 48}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return _SO


_lib = None


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint8))


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_bench_verify.restype = C.c_double
    return _lib


def ncores() -> int:
    """Host cores this process may actually use: CPU affinity capped by the cgroup CPU quota (a
    container that sees 128 CPUs but holds a 16-CPU quota gets throttled with 128 busy threads)."""
    n = len(os.sched_getaffinity(0))
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            with open(path) as f:
                parts = f.read().split()
            if path.endswith("cpu.max"):
                if parts[0] != "max":
                    n = min(n, max(1, int(int(parts[0]) / int(parts[1]))))
            else:
                q = int(parts[0])
                if q > 0:
                    with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as g:
                        n = min(n, max(1, q // int(g.read().split()[0])))
            break
        except (OSError, ValueError, IndexError):
            continue
    return n


def verify_batch(curve, r, s, qx, qy, digest, nthreads=None) -> np.ndarray:
    L = FIELD_BYTES[curve]
    r, pr = _u8(r); s, ps = _u8(s); qx, pqx = _u8(qx); qy, pqy = _u8(qy); digest, pd = _u8(digest)
    n = r.size // L
    dlen = digest.size // n if n else 32
    ok = np.zeros(n, dtype=np.uint8)
    lib().orc_verify_batch(C.c_int(curve), C.c_size_t(n), pr, ps, pqx, pqy, pd, C.c_size_t(dlen),
                           ok.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_int(nthreads or ncores()))
    return ok


def verify_batch_der(curve, sigs, sig_off, qxy, digest, nthreads=None) -> np.ndarray:
    sigs, psg = _u8(sigs if len(sigs) else np.zeros(1, np.uint8))
    sig_off = np.ascontiguousarray(sig_off, dtype=np.uint32)
    qxy, pq = _u8(qxy); digest, pd = _u8(digest)
    n = sig_off.size - 1
    dlen = digest.size // n if n else 32
    ok = np.zeros(n, dtype=np.uint8)
    lib().orc_verify_batch_der(C.c_int(curve), C.c_size_t(n), psg, sig_off.ctypes.data_as(C.POINTER(C.c_uint32)),
                               pq, pd, C.c_size_t(dlen), ok.ctypes.data_as(C.POINTER(C.c_uint8)),
                               C.c_int(nthreads or ncores()))
    return ok


def pubkey(curve, d: bytes):
    L = FIELD_BYTES[curve]
    qx = (C.c_uint8 * L)(); qy = (C.c_uint8 * L)()
    rc = lib().orc_pubkey(C.c_int(curve), (C.c_uint8 * L).from_buffer_copy(d), qx, qy)
    if rc:
        raise ValueError("bad private scalar")
    return bytes(qx), bytes(qy)


def sign_batch(curve, d_table, key_idx, digest, k):
    """Deterministic ECDSA signing with caller-supplied nonces (corpus generator)."""
    L = FIELD_BYTES[curve]
    d_table, pdt = _u8(d_table); digest, pd = _u8(digest); k, pk = _u8(k)
    key_idx = np.ascontiguousarray(key_idx, dtype=np.uint32)
    n = key_idx.size
    dlen = digest.size // n
    r = np.zeros((n, L), np.uint8); s = np.zeros((n, L), np.uint8)
    rc = lib().orc_sign_batch(C.c_int(curve), C.c_size_t(n), pdt, key_idx.ctypes.data_as(C.POINTER(C.c_uint32)),
                              pd, C.c_size_t(dlen), pk, r.ctypes.data_as(C.POINTER(C.c_uint8)),
                              s.ctypes.data_as(C.POINTER(C.c_uint8)))
    if rc < 0:
        raise RuntimeError("orc_sign_batch failed")
    return r, s


def lincomb(curve, a: bytes, b: bytes, qx: bytes, qy: bytes):
    """a*G + b*Q → (x, y) bytes or None for infinity."""
    L = FIELD_BYTES[curve]
    mk = lambda v: (C.c_uint8 * L).from_buffer_copy(v)
    ox = (C.c_uint8 * L)(); oy = (C.c_uint8 * L)()
    rc = lib().orc_lincomb(C.c_int(curve), mk(a), mk(b), mk(qx), mk(qy), ox, oy)
    if rc == 1:
        return None
    if rc:
        raise ValueError("orc_lincomb failed")
    return bytes(ox), bytes(oy)


def sha256_batch(msgs, off, nthreads=None) -> np.ndarray:
    msgs, pm = _u8(msgs if len(msgs) else np.zeros(1, np.uint8))
    off = np.ascontiguousarray(off, dtype=np.uint64)
    n = off.size - 1
    out = np.zeros((n, 32), np.uint8)
    lib().orc_sha256_batch(C.c_size_t(n), pm, off.ctypes.data_as(C.POINTER(C.c_uint64)),
                           out.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_int(nthreads or ncores()))
    return out


def bench_verify(curve, r, s, keys, key_idx, digest, nthreads=None):
    """OpenSSL ECDSA_do_verify over pre-built EC_KEYs; returns (seconds, verdicts)."""
    L = FIELD_BYTES[curve]
    r, pr = _u8(r); s, ps = _u8(s); keys, pk = _u8(keys); digest, pd = _u8(digest)
    key_idx = np.ascontiguousarray(key_idx, dtype=np.uint32)
    n = key_idx.size
    K = keys.size // (2 * L)
    dlen = digest.size // n
    ok = np.zeros(n, np.uint8)
    t = lib().orc_bench_verify(C.c_int(curve), C.c_size_t(n), pr, ps, pk, C.c_size_t(K),
                               key_idx.ctypes.data_as(C.POINTER(C.c_uint32)), pd, C.c_size_t(dlen),
                               ok.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_int(nthreads or ncores()))
    return float(t), ok
