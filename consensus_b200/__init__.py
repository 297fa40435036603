"""consensus_b200 — B200-native batched signature verification behind SmartBFT's api.Verifier.

The product is ``libsbv.so`` (hand-written sm_100a CUDA + a C ABI, include/sbv.h).  This package is
the thin ctypes binding used by the tests and bench.py; it never falls back to a CPU
implementation — if the library is missing or no CUDA device is usable, it raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsbv.so")

P256, P384 = 0, 1
FIELD_BYTES = {P256: 32, P384: 48}

SYMBOLS = [
    "sbv_create", "sbv_destroy", "sbv_last_error", "sbv_device_count", "sbv_verify_batch",
    "sbv_verify_batch_device", "sbv_verify_batch_der", "sbv_sha256_batch", "sbv_hash_verify_batch",
    "sbv_verify_mixed", "sbv_quorum", "sbv_compute_quorum", "sbv_set_keys", "sbv_kernel_launches",
    "sbv_probe_mad_rate", "sbv_profile_enable", "sbv_profile_read", "sbv_verify_registered",
    "sbv_verify_registered_device", "sbv_hash_verify_registered", "sbv_prepare_quorum", "sbv_verify_quorum",
    "sbv_comm_unique_id", "sbv_comm_init_rank", "sbv_comm_ranks", "sbv_gather_verdicts_device", "sbv_gather_words_device",
    "sbv_verify_batch_ranked", "sbv_host_alloc", "sbv_host_free",
]


class EngineFault(RuntimeError):
    """An engine fault (CUDA error, bad argument).  Never a verdict — callers must fail-stop."""


_lib = None


def load_library() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineFault(f"{LIB_PATH} is missing — run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        lib.sbv_last_error.restype = C.c_char_p
        lib.sbv_kernel_launches.restype = C.c_uint64
        lib.sbv_probe_mad_rate.restype = C.c_double
        lib.sbv_destroy.restype = None
        lib.sbv_compute_quorum.restype = None
        _lib = lib
    return _lib


def _p8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _u8(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


def compute_quorum(n: int):
    q, f = C.c_uint32(), C.c_uint32()
    load_library().sbv_compute_quorum(C.c_uint64(n), C.byref(q), C.byref(f))
    return q.value, f.value


class Engine:
    """One engine = 1..8 GPUs of one box (sbv_create)."""

    def __init__(self, devices=None, n_devices: int = 1):
        self._lib = load_library()
        self._h = C.c_void_p()
        if devices is not None:
            n_devices = len(devices)
            arr = (C.c_int * n_devices)(*devices)
        else:
            arr = None
        rc = self._lib.sbv_create(arr, C.c_int(n_devices), C.byref(self._h))
        if rc != 0:
            raise EngineFault(f"sbv_create failed ({rc}): no usable CUDA device? (there is no CPU fallback)")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.sbv_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc, what):
        if rc != 0:
            raise EngineFault(f"{what} failed ({rc}): {self._lib.sbv_last_error(self._h).decode()}")

    @property
    def device_count(self) -> int:
        return self._lib.sbv_device_count(self._h)

    @property
    def kernel_launches(self) -> int:
        return int(self._lib.sbv_kernel_launches(self._h))

    def probe_mad_rate(self) -> float:
        return float(self._lib.sbv_probe_mad_rate(self._h))

    def profile_enable(self, on=True):
        self._check(self._lib.sbv_profile_enable(self._h, C.c_int(1 if on else 0)), "sbv_profile_enable")

    def profile_read(self):
        """(prep_ms, verify_ms, n_launch_pairs) summed since the last read; synchronise first."""
        p, v, k = C.c_double(), C.c_double(), C.c_uint64()
        self._check(self._lib.sbv_profile_read(self._h, C.byref(p), C.byref(v), C.byref(k)), "sbv_profile_read")
        return p.value, v.value, k.value

    # ---- host-buffer API (numpy arrays, or anything exposing a host pointer via .ctypes) ----
    def verify_batch(self, curve, r, s, qx, qy, digest, out=None) -> np.ndarray:
        L = FIELD_BYTES[curve]
        r, s, qx, qy, digest = map(_u8, (r, s, qx, qy, digest))
        n = r.size // L
        dlen = digest.size // n if n else 32
        ok = out if out is not None else np.zeros(n, np.uint8)
        self._check(self._lib.sbv_verify_batch(self._h, C.c_uint8(curve), C.c_size_t(n), _p8(r), _p8(s), _p8(qx), _p8(qy),
                                               _p8(digest), C.c_uint8(dlen), _p8(ok)), "sbv_verify_batch")
        return ok

    def verify_batch_ptr(self, curve, n, r, s, qx, qy, digest, dlen, ok):
        """Raw host pointers (ints) — used with pinned torch tensors."""
        vp = C.c_void_p
        self._check(self._lib.sbv_verify_batch(self._h, C.c_uint8(curve), C.c_size_t(n), vp(r), vp(s), vp(qx), vp(qy),
                                               vp(digest), C.c_uint8(dlen), vp(ok)), "sbv_verify_batch")

    def verify_batch_device(self, curve, n, d_r, d_s, d_qx, d_qy, d_digest, dlen, d_ok, stream=0, device_index=0):
        """Device pointers (ints); enqueued on `stream` (cudaStream_t as int), not synchronised."""
        vp = C.c_void_p
        self._check(self._lib.sbv_verify_batch_device(self._h, C.c_int(device_index), C.c_uint8(curve), C.c_size_t(n), vp(d_r),
                                                      vp(d_s), vp(d_qx), vp(d_qy), vp(d_digest), C.c_uint8(dlen), vp(d_ok),
                                                      vp(stream)), "sbv_verify_batch_device")

    def set_keys(self, curves, xy, ids=None, verification_seq=0):
        """Registers keys (slot i = key i) and builds their comb tables.  xy: (n, 2, L) or (n, 96) bytes."""
        curves = _u8(curves)
        n = curves.size
        xy = np.asarray(xy, dtype=np.uint8)
        if xy.size != n * 96:  # fixed-width per-curve arrays -> 48-byte slots
            L = xy.size // (2 * n)
            slots = np.zeros((n, 2, 48), np.uint8)
            slots[:, :, 48 - L:] = xy.reshape(n, 2, L)
            xy = slots
        xy = np.ascontiguousarray(xy.reshape(-1))
        ids = np.ascontiguousarray(ids if ids is not None else np.arange(n), dtype=np.uint64)
        self._check(self._lib.sbv_set_keys(self._h, C.c_uint64(verification_seq), C.c_size_t(n), ids.ctypes.data_as(C.POINTER(C.c_uint64)),
                                           _p8(curves), _p8(xy)), "sbv_set_keys")

    def verify_registered(self, curve, key_slot, r, s, digest) -> np.ndarray:
        key_slot = np.ascontiguousarray(key_slot, dtype=np.uint32)
        r, s, digest = map(_u8, (r, s, digest))
        n = key_slot.size
        dlen = digest.size // n if n else 32
        ok = np.zeros(n, np.uint8)
        self._check(self._lib.sbv_verify_registered(self._h, C.c_uint8(curve), C.c_size_t(n), key_slot.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                    _p8(r), _p8(s), _p8(digest), C.c_uint8(dlen), _p8(ok)), "sbv_verify_registered")
        return ok

    def hash_verify_registered(self, curve, msgs, off, key_slot, r, s) -> np.ndarray:
        msgs = _u8(msgs if len(msgs) else np.zeros(1, np.uint8))
        off = np.ascontiguousarray(off, dtype=np.uint64)
        key_slot = np.ascontiguousarray(key_slot, dtype=np.uint32)
        r, s = _u8(r), _u8(s)
        n = off.size - 1
        ok = np.zeros(n, np.uint8)
        self._check(self._lib.sbv_hash_verify_registered(self._h, C.c_uint8(curve), C.c_size_t(n), _p8(msgs), off.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                         key_slot.ctypes.data_as(C.POINTER(C.c_uint32)), _p8(r), _p8(s), _p8(ok)),
                    "sbv_hash_verify_registered")
        return ok

    def verify_registered_device(self, curve, n, d_slot, d_r, d_s, d_digest, dlen, d_ok, stream=0, device_index=0):
        vp = C.c_void_p
        self._check(self._lib.sbv_verify_registered_device(self._h, C.c_int(device_index), C.c_uint8(curve), C.c_size_t(n), vp(d_slot), vp(d_r),
                                                           vp(d_s), vp(d_digest), C.c_uint8(dlen), vp(d_ok), vp(stream)),
                    "sbv_verify_registered_device")

    def verify_batch_der(self, curve, sigs, sig_off, qxy, digest) -> np.ndarray:
        sigs = _u8(sigs if len(sigs) else np.zeros(1, np.uint8))
        sig_off = np.ascontiguousarray(sig_off, dtype=np.uint32)
        qxy, digest = _u8(qxy), _u8(digest)
        n = sig_off.size - 1
        dlen = digest.size // n if n else 32
        ok = np.zeros(n, np.uint8)
        self._check(self._lib.sbv_verify_batch_der(self._h, C.c_uint8(curve), C.c_size_t(n), _p8(sigs),
                                                   sig_off.ctypes.data_as(C.POINTER(C.c_uint32)), _p8(qxy), _p8(digest),
                                                   C.c_uint8(dlen), _p8(ok)), "sbv_verify_batch_der")
        return ok

    def sha256_batch(self, msgs, off) -> np.ndarray:
        msgs = _u8(msgs if len(msgs) else np.zeros(1, np.uint8))
        off = np.ascontiguousarray(off, dtype=np.uint64)
        n = off.size - 1
        out = np.zeros((n, 32), np.uint8)
        self._check(self._lib.sbv_sha256_batch(self._h, C.c_size_t(n), _p8(msgs), off.ctypes.data_as(C.POINTER(C.c_uint64)),
                                               _p8(out)), "sbv_sha256_batch")
        return out

    def hash_verify_batch(self, curve, msgs, off, r, s, qx, qy, want_digest=False):
        msgs = _u8(msgs if len(msgs) else np.zeros(1, np.uint8))
        off = np.ascontiguousarray(off, dtype=np.uint64)
        r, s, qx, qy = map(_u8, (r, s, qx, qy))
        n = off.size - 1
        ok = np.zeros(n, np.uint8)
        dig = np.zeros((n, 32), np.uint8) if want_digest else None
        self._check(self._lib.sbv_hash_verify_batch(self._h, C.c_uint8(curve), C.c_size_t(n), _p8(msgs),
                                                    off.ctypes.data_as(C.POINTER(C.c_uint64)), _p8(r), _p8(s), _p8(qx), _p8(qy),
                                                    _p8(dig) if want_digest else None, _p8(ok)), "sbv_hash_verify_batch")
        return (ok, dig) if want_digest else ok

    def verify_mixed(self, curve_tag, r48, s48, qx48, qy48, digest32) -> np.ndarray:
        curve_tag, r48, s48, qx48, qy48, digest32 = map(_u8, (curve_tag, r48, s48, qx48, qy48, digest32))
        n = curve_tag.size
        ok = np.zeros(n, np.uint8)
        self._check(self._lib.sbv_verify_mixed(self._h, C.c_size_t(n), _p8(curve_tag), _p8(r48), _p8(s48), _p8(qx48), _p8(qy48),
                                               _p8(digest32), _p8(ok)), "sbv_verify_mixed")
        return ok

    # ---- one process per GPU ----
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        rc = load_library().sbv_comm_unique_id(buf)
        if rc != 0:
            raise EngineFault(f"sbv_comm_unique_id failed ({rc})")
        return bytes(buf)

    def comm_init_rank(self, uid: bytes, nranks: int, rank: int) -> int:
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        ch = self._lib.sbv_comm_init_rank(self._h, buf, C.c_int(nranks), C.c_int(rank))
        if ch < 0:
            self._check(ch, "sbv_comm_init_rank")
        return ch

    def gather_verdicts_device(self, channel, d_ok, n, d_mask_all, stream=0):
        vp = C.c_void_p
        self._check(self._lib.sbv_gather_verdicts_device(self._h, C.c_int(channel), vp(d_ok), C.c_size_t(n), vp(d_mask_all), vp(stream)),
                    "sbv_gather_verdicts_device")

    def gather_words_device(self, channel, d_all, words, stream=0):
        vp = C.c_void_p
        self._check(self._lib.sbv_gather_words_device(self._h, C.c_int(channel), vp(d_all), C.c_size_t(words), vp(stream)), "sbv_gather_words_device")

    def verify_batch_ranked_ptr(self, channel, curve, n, r, s, qx, qy, digest, dlen, ok, mask_all):
        vp = C.c_void_p
        self._check(self._lib.sbv_verify_batch_ranked(self._h, C.c_int(channel), C.c_uint8(curve), C.c_size_t(n), vp(r), vp(s), vp(qx), vp(qy),
                                                      vp(digest), C.c_uint8(dlen), vp(ok), vp(mask_all)), "sbv_verify_batch_ranked")

    def verify_quorum(self, curve, r, s, qx, qy, digest, instance, sender, signer, digest_match, n_instances, threshold, self_id=None):
        """Commit votes: signatures verified, verdicts counted on the device.  Returns (ok, valid_count, reached)."""
        r, s, qx, qy, digest, digest_match = map(_u8, (r, s, qx, qy, digest, digest_match))
        instance = np.ascontiguousarray(instance, dtype=np.uint32)
        sender = np.ascontiguousarray(sender, dtype=np.uint16)
        signer = np.ascontiguousarray(signer, dtype=np.uint16)
        n = instance.size
        dlen = digest.size // n if n else 32
        ok = np.zeros(n, np.uint8)
        cnt = np.zeros(n_instances, np.uint32)
        reached = np.zeros(n_instances, np.uint8)
        sid = None
        if self_id is not None:
            self_id = np.ascontiguousarray(self_id, dtype=np.uint16)
            sid = self_id.ctypes.data_as(C.POINTER(C.c_uint16))
        u16 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint16))
        self._check(self._lib.sbv_verify_quorum(self._h, C.c_uint8(curve), C.c_size_t(n), _p8(r), _p8(s), _p8(qx), _p8(qy), _p8(digest), C.c_uint8(dlen),
                                                instance.ctypes.data_as(C.POINTER(C.c_uint32)), u16(sender), u16(signer), _p8(digest_match),
                                                C.c_size_t(n_instances), sid, C.c_uint32(threshold), _p8(ok),
                                                cnt.ctypes.data_as(C.POINTER(C.c_uint32)), _p8(reached)), "sbv_verify_quorum")
        return ok, cnt, reached

    def prepare_quorum(self, instance, sender, digest_match, n_instances, threshold, self_id=None):
        instance = np.ascontiguousarray(instance, dtype=np.uint32)
        sender = np.ascontiguousarray(sender, dtype=np.uint16)
        digest_match = _u8(digest_match)
        cnt = np.zeros(n_instances, np.uint32)
        reached = np.zeros(n_instances, np.uint8)
        sid = None
        if self_id is not None:
            self_id = np.ascontiguousarray(self_id, dtype=np.uint16)
            sid = self_id.ctypes.data_as(C.POINTER(C.c_uint16))
        self._check(self._lib.sbv_prepare_quorum(self._h, C.c_size_t(instance.size), instance.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                 sender.ctypes.data_as(C.POINTER(C.c_uint16)), _p8(digest_match), C.c_size_t(n_instances), sid,
                                                 C.c_uint32(threshold), cnt.ctypes.data_as(C.POINTER(C.c_uint32)), _p8(reached)), "sbv_prepare_quorum")
        return cnt, reached

    def quorum(self, instance, sender, signer, digest_match, ok, n_instances, threshold, self_id=None):
        instance = np.ascontiguousarray(instance, dtype=np.uint32)
        sender = np.ascontiguousarray(sender, dtype=np.uint16)
        signer = np.ascontiguousarray(signer, dtype=np.uint16)
        digest_match, ok = _u8(digest_match), _u8(ok)
        cnt = np.zeros(n_instances, np.uint32)
        reached = np.zeros(n_instances, np.uint8)
        sid = None
        if self_id is not None:
            self_id = np.ascontiguousarray(self_id, dtype=np.uint16)
            sid = self_id.ctypes.data_as(C.POINTER(C.c_uint16))
        self._check(self._lib.sbv_quorum(self._h, C.c_size_t(instance.size), instance.ctypes.data_as(C.POINTER(C.c_uint32)),
                                         sender.ctypes.data_as(C.POINTER(C.c_uint16)), signer.ctypes.data_as(C.POINTER(C.c_uint16)),
                                         _p8(digest_match), _p8(ok), C.c_size_t(n_instances), sid, C.c_uint32(threshold),
                                         cnt.ctypes.data_as(C.POINTER(C.c_uint32)), _p8(reached)), "sbv_quorum")
        return cnt, reached
