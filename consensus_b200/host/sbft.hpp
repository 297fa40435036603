// sbft.hpp — host-side value types and pure helpers mirrored from the reference (C++17, header only).
//
//   Proposal, Signature, RequestInfo          /root/reference/pkg/types/types.go:18-44
//   Proposal::Digest                          pkg/types/types.go:50-69   (hex SHA-256 of asn1.Marshal)
//   CommitSignaturesDigest                    internal/bft/util.go:564-595
//   computeQuorum                             internal/bft/util.go:183-187
//   PreparesFrom / ViewMetadata wire codecs   smartbftprotos/messages.proto:56-58, 105-111
//
// Hashing ONE message is a sequential chain (the reference does it on one core, types.go:64-69), so
// the single-proposal digest is computed on the host; batches of messages go to the GPU through
// sbv_sha256_batch / sbv_hash_verify_batch (verifier.hpp).
#pragma once
#include <cstdint>
#include <cstring>
#include <optional>
#include <string>
#include <vector>

namespace sbft {

using Bytes = std::vector<uint8_t>;
// Go's `error`: nullopt == nil
using Error = std::optional<std::string>;
inline Error Errorf(std::string s) { return Error(std::move(s)); }

// ---------------------------------------------------------------- SHA-256 (FIPS 180-4), host
namespace detail {
inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
inline void compress(uint32_t h[8], const uint8_t *blk) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++) w[i] = (uint32_t)blk[4 * i] << 24 | (uint32_t)blk[4 * i + 1] << 16 | (uint32_t)blk[4 * i + 2] << 8 | blk[4 * i + 3];
    for (int i = 16; i < 64; i++) {
        uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
        uint32_t t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + K256[i] + w[i];
        uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
}  // namespace detail

inline Bytes sha256(const uint8_t *p, size_t n) {
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    size_t i = 0;
    for (; i + 64 <= n; i += 64) detail::compress(h, p + i);
    uint8_t tail[128] = {0};
    size_t rem = n - i;
    if (rem) memcpy(tail, p + i, rem);
    tail[rem] = 0x80;
    size_t tl = rem + 9 <= 64 ? 64 : 128;
    uint64_t bits = (uint64_t)n * 8;
    for (int k = 0; k < 8; k++) tail[tl - 1 - k] = (uint8_t)(bits >> (8 * k));
    detail::compress(h, tail);
    if (tl == 128) detail::compress(h, tail + 64);
    Bytes out(32);
    for (int k = 0; k < 8; k++) { out[4 * k] = h[k] >> 24; out[4 * k + 1] = h[k] >> 16; out[4 * k + 2] = h[k] >> 8; out[4 * k + 3] = h[k]; }
    return out;
}
inline Bytes sha256(const Bytes &b) { return sha256(b.data(), b.size()); }
inline std::string hex(const Bytes &b) {
    static const char *d = "0123456789abcdef";
    std::string s;
    for (uint8_t c : b) { s.push_back(d[c >> 4]); s.push_back(d[c & 15]); }
    return s;
}

// ---------------------------------------------------------------- encoding/asn1 Marshal restatement
namespace der {
inline void len(Bytes &o, size_t n) {
    if (n < 0x80) { o.push_back((uint8_t)n); return; }
    uint8_t tmp[8]; int k = 0;
    while (n) { tmp[k++] = (uint8_t)n; n >>= 8; }
    o.push_back(0x80 | k);
    while (k) o.push_back(tmp[--k]);
}
inline void octets(Bytes &o, const Bytes &b) { o.push_back(0x04); len(o, b.size()); o.insert(o.end(), b.begin(), b.end()); }
inline void int64(Bytes &o, int64_t v) {  // minimal two's complement
    int n = 1;
    while (n < 8 && !(v >= -(int64_t(1) << (8 * n - 1)) && v < (int64_t(1) << (8 * n - 1)))) n++;
    o.push_back(0x02); o.push_back((uint8_t)n);
    for (int k = n - 1; k >= 0; k--) o.push_back((uint8_t)((uint64_t)v >> (8 * k)));
}
inline Bytes seq(const Bytes &body) { Bytes o{0x30}; len(o, body.size()); o.insert(o.end(), body.begin(), body.end()); return o; }
}  // namespace der

// ---------------------------------------------------------------- types
struct Proposal {  // types.go:18-23 — asn1 field order is the struct order
    Bytes Payload, Header, Metadata;
    int64_t VerificationSequence = 0;
    Bytes Der() const {
        Bytes body;
        der::octets(body, Payload); der::octets(body, Header); der::octets(body, Metadata); der::int64(body, VerificationSequence);
        return der::seq(body);
    }
    Bytes DigestRaw() const { return sha256(Der()); }
    std::string Digest() const { return hex(DigestRaw()); }  // types.go:50-69
    bool operator==(const Proposal &o) const {
        return Payload == o.Payload && Header == o.Header && Metadata == o.Metadata && VerificationSequence == o.VerificationSequence;
    }
};
struct Signature {  // types.go:25-29
    uint64_t ID = 0;
    Bytes Value, Msg;
};
struct RequestInfo {  // types.go:41-44
    std::string ClientID, ID;
    bool operator==(const RequestInfo &o) const { return ClientID == o.ClientID && ID == o.ID; }
};

// CommitSignaturesDigest — util.go:564-586; empty input -> empty (nil)
inline Bytes CommitSignaturesDer(const std::vector<Signature> &sigs) {  // asn1.Marshal(IntDoubleBytes{A: [...]}) — util.go:570-578
    Bytes inner;
    for (const auto &s : sigs) {
        Bytes one;
        der::int64(one, (int64_t)s.ID); der::octets(one, s.Value); der::octets(one, s.Msg);
        Bytes sq = der::seq(one);
        inner.insert(inner.end(), sq.begin(), sq.end());
    }
    return der::seq(der::seq(inner));
}
inline Bytes CommitSignaturesDigest(const std::vector<Signature> &sigs) {
    if (sigs.empty()) return {};
    return sha256(CommitSignaturesDer(sigs));
}

// computeQuorum — util.go:183-187 (ceil((n+f+1)/2) == (n+f+2)/2 in integers)
inline void computeQuorum(uint64_t n, int &q, int &f) {
    f = n ? (int)((n - 1) / 3) : 0;
    q = (int)((n + f + 2) / 2);
}

// ---------------------------------------------------------------- minimal protobuf codecs
namespace pb {
inline void varint(Bytes &o, uint64_t v) { while (v >= 0x80) { o.push_back((uint8_t)v | 0x80); v >>= 7; } o.push_back((uint8_t)v); }
inline bool read_varint(const uint8_t *&p, const uint8_t *end, uint64_t &v) {
    v = 0;
    for (int sh = 0; sh < 64 && p < end; sh += 7) { uint8_t b = *p++; v |= (uint64_t)(b & 0x7f) << sh; if (!(b & 0x80)) return true; }
    return false;
}
// skips one field of wire type wt; false on malformed input
inline bool skip(const uint8_t *&p, const uint8_t *end, unsigned wt) {
    uint64_t v;
    switch (wt) {
        case 0: return read_varint(p, end, v);
        case 1: if (end - p < 8) return false; p += 8; return true;
        case 2: if (!read_varint(p, end, v) || (uint64_t)(end - p) < v) return false; p += v; return true;
        case 5: if (end - p < 4) return false; p += 4; return true;
        default: return false;
    }
}
}  // namespace pb

struct PreparesFrom {  // messages.proto:56-58
    std::vector<uint64_t> Ids;
    Bytes Marshal() const {  // proto3 packed encoding, as golang/protobuf emits
        Bytes o;
        if (Ids.empty()) return o;
        Bytes packed;
        for (uint64_t id : Ids) pb::varint(packed, id);
        o.push_back(0x0A); pb::varint(o, packed.size()); o.insert(o.end(), packed.begin(), packed.end());
        return o;
    }
    static bool Unmarshal(const Bytes &b, PreparesFrom &out) {
        out.Ids.clear();
        const uint8_t *p = b.data(), *end = p + b.size();
        while (p < end) {
            uint64_t key;
            if (!pb::read_varint(p, end, key)) return false;
            unsigned field = (unsigned)(key >> 3), wt = (unsigned)(key & 7);
            if (field == 0) return false;
            if (field == 1 && wt == 2) {
                uint64_t l;
                if (!pb::read_varint(p, end, l) || (uint64_t)(end - p) < l) return false;
                const uint8_t *q = p, *qe = p + l;
                while (q < qe) { uint64_t v; if (!pb::read_varint(q, qe, v)) return false; out.Ids.push_back(v); }
                p = qe;
            } else if (field == 1 && wt == 0) {
                uint64_t v; if (!pb::read_varint(p, end, v)) return false; out.Ids.push_back(v);
            } else if (!pb::skip(p, end, wt)) return false;
        }
        return true;
    }
};

struct ViewMetadata {  // messages.proto:105-111 (fields 1..3 are what the restated call sites read)
    uint64_t ViewId = 0, LatestSequence = 0, DecisionsInView = 0;
    Bytes Marshal() const {
        Bytes o;
        if (ViewId) { o.push_back(0x08); pb::varint(o, ViewId); }
        if (LatestSequence) { o.push_back(0x10); pb::varint(o, LatestSequence); }
        if (DecisionsInView) { o.push_back(0x18); pb::varint(o, DecisionsInView); }
        return o;
    }
    static bool Unmarshal(const Bytes &b, ViewMetadata &out) {
        out = ViewMetadata();
        const uint8_t *p = b.data(), *end = p + b.size();
        while (p < end) {
            uint64_t key;
            if (!pb::read_varint(p, end, key)) return false;
            unsigned field = (unsigned)(key >> 3), wt = (unsigned)(key & 7);
            if (field == 0) return false;
            if (field <= 3 && wt == 0) {
                uint64_t v; if (!pb::read_varint(p, end, v)) return false;
                (field == 1 ? out.ViewId : field == 2 ? out.LatestSequence : out.DecisionsInView) = v;
            } else if (!pb::skip(p, end, wt)) return false;
        }
        return true;
    }
};

}  // namespace sbft
