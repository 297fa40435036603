"""oracle/ecdsa_ref.py — pure-Python CPU ORACLE (big-int restatement).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module;
the product path (consensus_b200 / libsbv.so) never does.

Independent of OpenSSL: affine big-integer arithmetic only.  It restates

* ECDSA verification with the Go ``crypto/ecdsa.VerifyASN1`` accept set (Go stdlib, pinned by the
  reference's go.mod:3 "go 1.20" / CI go 1.21.8; absent from /root/reference and from this image,
  so the published algorithm — FIPS 186-4 §6.4 — is restated and pinned on RFC 6979 A.2.5/A.2.6);
* ``types.Proposal.Digest``          — /root/reference/pkg/types/types.go:50-69
* ``CommitSignaturesDigest``         — internal/bft/util.go:564-595
* ``computeQuorum``                  — internal/bft/util.go:183-187 (golden table util_test.go:144-154)
* the commit-vote acceptance rule    — internal/bft/view.go:161-171 (Signer == sender),
  util.go:130-143 (one vote per sender), view.go:827-849 (digest match, then VerifyConsenterSig),
  view.go:531 (Quorum-1 valid foreign votes)
* ``ValidateLastDecision`` counting  — internal/bft/viewchanger.go:697-727

PARITY PINNING: the reference has no golden vector for any of these values ("parity unpinned" at
reference level); see tests/test_oracle.py for what this oracle is pinned against.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass

P256, P384 = 0, 1


@dataclass(frozen=True)
class Curve:
    name: str
    p: int
    a: int
    b: int
    n: int
    gx: int
    gy: int
    size: int  # bytes per field element / scalar


CURVES = {
    P256: Curve(
        "P-256",
        0xFFFFFFFF00000001000000000000000000000000FFFFFFFFFFFFFFFFFFFFFFFF,
        -3,
        0x5AC635D8AA3A93E7B3EBBD55769886BC651D06B0CC53B0F63BCE3C3E27D2604B,
        0xFFFFFFFF00000000FFFFFFFFFFFFFFFFBCE6FAADA7179E84F3B9CAC2FC632551,
        0x6B17D1F2E12C4247F8BCE6E563A440F277037D812DEB33A0F4A13945D898C296,
        0x4FE342E2FE1A7F9B8EE7EB4A7C0F9E162BCE33576B315ECECBB6406837BF51F5,
        32,
    ),
    P384: Curve(
        "P-384",
        0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEFFFFFFFF0000000000000000FFFFFFFF,
        -3,
        0xB3312FA7E23EE7E4988E056BE3F82D19181D9C6EFE8141120314088F5013875AC656398D8A2ED19D2A85C8EDD3EC2AEF,
        0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFC7634D81F4372DDF581A0DB248B0A77AECEC196ACCC52973,
        0xAA87CA22BE8B05378EB1C71EF320AD746E1D3B628BA79B9859F741E082542A385502F25DBF55296C3A545E3872760AB7,
        0x3617DE4A96262C6F5D9E98BF9292DC29F8F41DBD289A147CE9DA3113B5F0B8C00A60B1CE1D7E819D7A431D7C90EA0E5F,
        48,
    ),
}


# ---------------------------------------------------------------- affine group law (None = infinity)
def _add(c: Curve, P, Q):
    if P is None:
        return Q
    if Q is None:
        return P
    x1, y1 = P
    x2, y2 = Q
    if x1 == x2:
        if (y1 + y2) % c.p == 0:
            return None
        lam = (3 * x1 * x1 + c.a) * pow(2 * y1, -1, c.p) % c.p
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, c.p) % c.p
    x3 = (lam * lam - x1 - x2) % c.p
    return x3, (lam * (x1 - x3) - y1) % c.p


def scalar_mult(c: Curve, k: int, P):
    R = None
    while k:
        if k & 1:
            R = _add(c, R, P)
        P = _add(c, P, P)
        k >>= 1
    return R


def on_curve(c: Curve, x: int, y: int) -> bool:
    return 0 <= x < c.p and 0 <= y < c.p and (y * y - (x * x * x + c.a * x + c.b)) % c.p == 0


def hash_to_int(c: Curve, digest: bytes) -> int:
    """Leftmost min(len, size) bytes as an integer (crypto/ecdsa hashToNat; no bit shift for P-256/384)."""
    return int.from_bytes(digest[: c.size], "big")


def verify(curve: int, qx: int, qy: int, digest: bytes, r: int, s: int) -> bool:
    c = CURVES[curve]
    if not on_curve(c, qx, qy):
        return False
    if not (1 <= r < c.n and 1 <= s < c.n):
        return False
    e = hash_to_int(c, digest)
    w = pow(s, -1, c.n)
    u1, u2 = e * w % c.n, r * w % c.n
    R = _add(c, scalar_mult(c, u1, (c.gx, c.gy)), scalar_mult(c, u2, (qx, qy)))
    if R is None:
        return False
    return R[0] % c.n == r


def verify_bytes(curve: int, r: bytes, s: bytes, qx: bytes, qy: bytes, digest: bytes) -> bool:
    f = lambda b: int.from_bytes(b, "big")
    return verify(curve, f(qx), f(qy), digest, f(r), f(s))


def pubkey(curve: int, d: int):
    c = CURVES[curve]
    return scalar_mult(c, d, (c.gx, c.gy))


def sign(curve: int, d: int, digest: bytes, k: int):
    c = CURVES[curve]
    e = hash_to_int(c, digest)
    R = scalar_mult(c, k, (c.gx, c.gy))
    r = R[0] % c.n
    s = pow(k, -1, c.n) * (e + r * d) % c.n
    return r, s


# ---------------------------------------------------------------- strict DER (VerifyASN1 / cryptobyte rules)
def der_parse(sig: bytes):
    """SEQUENCE{INTEGER r, INTEGER s}; minimal, non-negative, no trailing bytes. None if malformed."""
    def integer(buf, pos):
        if pos + 2 > len(buf) or buf[pos] != 0x02:
            return None
        l = buf[pos + 1]
        pos += 2
        if l & 0x80 or l == 0 or pos + l > len(buf):
            return None
        v = buf[pos : pos + l]
        if v[0] & 0x80:
            return None
        if l > 1 and v[0] == 0 and not (v[1] & 0x80):
            return None
        return int.from_bytes(v, "big"), pos + l

    if len(sig) < 2 or sig[0] != 0x30:
        return None
    if sig[1] < 0x80:
        l, pos = sig[1], 2
    elif sig[1] == 0x81:
        if len(sig) < 3 or sig[2] < 0x80:
            return None
        l, pos = sig[2], 3
    else:
        return None
    if len(sig) - pos != l:
        return None
    a = integer(sig, pos)
    if a is None:
        return None
    r, pos = a
    b = integer(sig, pos)
    if b is None:
        return None
    s, pos = b
    if pos != len(sig):
        return None
    return r, s


def der_encode(r: int, s: int) -> bytes:
    def integer(v):
        b = v.to_bytes((v.bit_length() + 8) // 8 or 1, "big")  # leading 0 when top bit set
        return b"\x02" + _der_len(len(b)) + b
    body = integer(r) + integer(s)
    return b"\x30" + _der_len(len(body)) + body


def _der_len(n: int) -> bytes:
    if n < 0x80:
        return bytes([n])
    b = n.to_bytes((n.bit_length() + 7) // 8, "big")
    return bytes([0x80 | len(b)]) + b


def verify_der(curve: int, qx: int, qy: int, digest: bytes, sig: bytes) -> bool:
    rs = der_parse(sig)
    if rs is None:
        return False
    return verify(curve, qx, qy, digest, rs[0], rs[1])


# ---------------------------------------------------------------- encoding/asn1 Marshal restatement
def _der_octets(b: bytes) -> bytes:
    return b"\x04" + _der_len(len(b)) + b


def _der_int64(v: int) -> bytes:
    """encoding/asn1 INTEGER for an int64: minimal two's complement."""
    n = 1
    while not (-(1 << (8 * n - 1)) <= v < (1 << (8 * n - 1))):
        n += 1
    b = v.to_bytes(n, "big", signed=True)
    return b"\x02" + _der_len(len(b)) + b


def _der_seq(body: bytes) -> bytes:
    return b"\x30" + _der_len(len(body)) + body


def proposal_der(payload: bytes, header: bytes, metadata: bytes, verification_sequence: int) -> bytes:
    """asn1.Marshal(types.Proposal) — struct field order Payload, Header, Metadata, VerificationSequence
    (pkg/types/types.go:18-23)."""
    return _der_seq(_der_octets(payload) + _der_octets(header) + _der_octets(metadata) + _der_int64(verification_sequence))


def proposal_digest(payload: bytes, header: bytes, metadata: bytes, verification_sequence: int) -> str:
    """types.Proposal.Digest — pkg/types/types.go:50-69: hex(SHA-256(DER))."""
    return hashlib.sha256(proposal_der(payload, header, metadata, verification_sequence)).hexdigest()


def commit_signatures_der(sigs) -> bytes:
    """asn1.Marshal(IntDoubleBytes{A: []IntDoubleByte{A int64, B, C []byte}}) — util.go:588-595."""
    inner = b"".join(_der_seq(_der_int64(signer) + _der_octets(value) + _der_octets(msg)) for signer, value, msg in sigs)
    return _der_seq(_der_seq(inner))


def commit_signatures_digest(sigs):
    """CommitSignaturesDigest — internal/bft/util.go:564-586; None (nil) for empty input."""
    if not sigs:
        return None
    return hashlib.sha256(commit_signatures_der(sigs)).digest()


# ---------------------------------------------------------------- quorum rules
def compute_quorum(n: int):
    """computeQuorum — internal/bft/util.go:183-187.  ceil((n+f+1)/2) == (n+f+2)//2 in integers."""
    f = (n - 1) // 3
    q = (n + f + 2) // 2
    return q, f


def count_commit_votes(votes, self_id=None):
    """Distinct-signer valid-vote count for one instance.

    votes: iterable of (sender, signer, digest_match, sig_ok) in arrival order.
    view.go:161-171 — a commit vote is registered only if Signature.Signer == sender;
    util.go:135-141 — only the first registered vote of a sender counts (a later one is dropped even
    if the first turns out invalid); view.go:829-842 — it is valid iff digest matches and the
    signature verifies; view.go:214-217-ish sender == self never reaches the vote set.
    """
    voted = set()
    valid = 0
    for sender, signer, digest_match, sig_ok in votes:
        if self_id is not None and sender == self_id:
            continue
        if signer != sender:
            continue
        if sender in voted:
            continue
        voted.add(sender)
        if digest_match and sig_ok:
            valid += 1
    return valid


def count_commit_votes_batch(instance, sender, signer, digest_match, sig_ok, n_instances, threshold, self_id=None):
    """count_commit_votes over a stream of votes grouped by instance (arrival order inside an instance).
    Returns (valid_count[n_instances], reached[n_instances]) as Python lists -> numpy via the caller;
    reached = valid_count >= threshold (the caller passes Quorum-1, view.go:531).  self_id: per-instance ids or None."""
    import numpy as np
    cnt = np.zeros(n_instances, np.uint32)
    voted = {}
    inst_l, snd_l, sig_l = [int(x) for x in instance], [int(x) for x in sender], [int(x) for x in signer]
    dm_l, ok_l = [int(x) for x in digest_match], [int(x) for x in sig_ok]
    sid = None if self_id is None else [int(x) for x in self_id]
    for v in range(len(inst_l)):
        i, snd = inst_l[v], snd_l[v]
        if i >= n_instances:
            continue
        if sid is not None and snd == sid[i]:
            continue
        if sig_l[v] != snd:
            continue
        seen = voted.setdefault(i, set())
        if snd in seen:
            continue
        seen.add(snd)
        if dm_l[v] and ok_l[v]:
            cnt[i] += 1
    return cnt, (cnt >= threshold).astype(np.uint8)


def validate_last_decision_sigs(signers, sig_ok, quorum: int) -> bool:
    """viewchanger.go:697-727: >= quorum signatures present; duplicates of a signer skipped; any
    invalid (non-duplicate) signature fails the decision; valid distinct >= quorum."""
    if len(signers) < quorum:
        return False
    seen = set()
    valid = 0
    for sg, ok in zip(signers, sig_ok):
        if sg in seen:
            continue
        seen.add(sg)
        if not ok:
            return False
        valid += 1
    return valid >= quorum


# ---------------------------------------------------------------- deterministic byte source
class DRBG:
    """Counter-mode SHA-256: block i = SHA-256(ascii(seed) || b':' || ascii(i)). SURVEY.md §8d."""

    def __init__(self, seed: int):
        self.prefix = str(seed).encode() + b":"

    def block(self, i: int) -> bytes:
        return hashlib.sha256(self.prefix + str(i).encode()).digest()

    def bytes(self, i: int, n: int) -> bytes:
        out = b""
        j = 0
        while len(out) < n:
            out += hashlib.sha256(self.prefix + str(i).encode() + b"/" + str(j).encode()).digest()
            j += 1
        return out[:n]

    def u32(self, i: int) -> int:
        return int.from_bytes(self.block(i)[:4], "big")
