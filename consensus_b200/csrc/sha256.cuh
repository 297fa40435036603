// sha256.cuh — SHA-256 over a ragged batch, one message per thread (FIPS 180-4).
//
// Replaces the reference's crypto/sha256 call sites on the verification path:
// /root/reference/pkg/types/types.go:64-69 (computeDigest), internal/bft/util.go:583-585.
// Messages are concatenated in one device buffer with byte offsets off[n+1]; a thread walks its
// message with ALIGNED 32-bit loads and re-aligns with PRMT (byte_perm), so arbitrary byte
// offsets cost no byte loads.  The buffer must be readable 8 bytes past the last message.
#pragma once
#include <stdint.h>

namespace sbv {

__constant__ uint32_t SHA256_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

__device__ __forceinline__ uint32_t rotr(uint32_t x, int n) { return __funnelshift_r(x, x, n); }

__device__ __forceinline__ void sha256_compress(uint32_t (&h)[8], uint32_t (&w)[16]) {
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        if (i >= 16) {
            uint32_t w15 = w[(i - 15) & 15], w2 = w[(i - 2) & 15];
            uint32_t s0 = rotr(w15, 7) ^ rotr(w15, 18) ^ (w15 >> 3);
            uint32_t s1 = rotr(w2, 17) ^ rotr(w2, 19) ^ (w2 >> 10);
            w[i & 15] = w[i & 15] + s0 + w[(i - 7) & 15] + s1;
        }
        uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = hh + S1 + ch + SHA256_K[i] + w[i & 15];
        uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

// digest_out: 32 bytes per message, big-endian words (the byte string SHA-256 defines)
__global__ void __launch_bounds__(128) k_sha256(uint32_t n, const uint8_t *__restrict__ msgs,
                                                const uint64_t *__restrict__ off, uint64_t base, uint8_t *__restrict__ digest_out) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const uint64_t o = off[idx] - base;
    const uint64_t len = off[idx + 1] - off[idx];
    const uint32_t *words = reinterpret_cast<const uint32_t *>(msgs + (o & ~(uint64_t)3));
    const uint32_t sh = (uint32_t)(o & 3);
    const uint32_t sel = (sh + 3) | ((sh + 2) << 4) | ((sh + 1) << 8) | (sh << 12);
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    const uint64_t nblocks = (len + 9 + 63) / 64;
    for (uint64_t blk = 0; blk < nblocks; blk++) {
        uint32_t w[16];
        const uint64_t base = blk * 64;
        if (base + 64 <= len) {
            uint32_t prev = __ldg(words + blk * 16);
#pragma unroll
            for (int j = 0; j < 16; j++) {
                uint32_t next = (sh || j < 15) ? __ldg(words + blk * 16 + j + 1) : 0u;
                w[j] = __byte_perm(prev, next, sel);
                prev = next;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const uint64_t p = base + 4 * (uint64_t)j;
                uint32_t v = 0;
                if (p < len) {
                    uint32_t a = __ldg(words + blk * 16 + j), b = __ldg(words + blk * 16 + j + 1);
                    v = __byte_perm(a, b, sel);
                    uint32_t rem = (uint32_t)(len - p);  // valid bytes in this word (>= 1)
                    if (rem < 4) v = (v & (0xffffffffu << (8 * (4 - rem)))) | (0x80u << (8 * (3 - rem)));
                } else if (p == len) {
                    v = 0x80000000u;
                }
                w[j] = v;
            }
            if (blk == nblocks - 1) {
                const uint64_t bits = len * 8;
                w[14] = (uint32_t)(bits >> 32);
                w[15] = (uint32_t)bits;
            }
        }
        sha256_compress(h, w);
    }
    uint4 *out = reinterpret_cast<uint4 *>(digest_out + (size_t)idx * 32);
    out[0] = make_uint4(__byte_perm(h[0], 0, 0x0123), __byte_perm(h[1], 0, 0x0123), __byte_perm(h[2], 0, 0x0123), __byte_perm(h[3], 0, 0x0123));
    out[1] = make_uint4(__byte_perm(h[4], 0, 0x0123), __byte_perm(h[5], 0, 0x0123), __byte_perm(h[6], 0, 0x0123), __byte_perm(h[7], 0, 0x0123));
}

}  // namespace sbv
