// keygroup.cuh — per-key fixed-base tables, built on the device, and the grouping of a keys-per-item batch by key.
//
// In the reference a Verifier sees the same few keys over and over: the n consenters sign every commit vote
// (/root/reference/internal/bft/view.go:519-551: up to n-1 votes per sequence from the same n-1 nodes) and clients
// sign many requests each (controller.go:233-246).  The C ABI still takes the key with every item (sbv_verify_batch:
// qx, qy per item), so the engine finds the repetition itself:
//
//   k_kg_insert   every item hashes its 64/96-byte key into an open-addressing table (CAS on the item index,
//                 full-key compare on collision): rep[i] = first item with the same key; warp-aggregated count
//   k_kg_assign   representatives whose key occurs >= T times (and while table slots last) get a dense key id
//   k_kg_route    items are appended to the fixed-base list (their key has a table) or to the generic list
//   k_kt_bases4   four lanes per key: validate the key, B_w = 2^(W*w) * Q for all windows — a chain of doublings whose
//                 independent multiplications run on different lanes (k_kt_bases: the one-thread-per-key form)
//   k_kt_fill     one thread per (key, window): e*B_w for e = 1..2^(W-1) with co-Z additions (5M+2S each — the
//                 chain of Z ratios that comes with them is exactly what the inversion needs)
//   k_kt_inv      one thread per key: ONE field inversion for all windows of the key (Montgomery's trick across
//                 the windows' top Z's)
//   k_kt_final    one thread per (key, window): back-substitute the Z ratios, convert to affine, in place
//
// The table feeds k_verify_kt (kernels.cuh).  sbv_set_keys uses the same builder once per registration.
// Cost per key (P-256, W = 5: 52 windows x 16 entries): 255 doublings + 832 x ~12 multiplications + one inversion
// ~ 4 generic verifications; a fixed-base verification is ~4.5x cheaper than a generic one, so T = 16 pays.
#pragma once
#include "kernels.cuh"

namespace sbv {

constexpr uint32_t KG_EMPTY = 0xffffffffu;

template <class C>
SBV_DEV uint32_t kg_hash(const uint8_t *__restrict__ qx_be, const uint8_t *__restrict__ qy_be, uint32_t i, uint32_t seed) {
    const uint32_t *x = reinterpret_cast<const uint32_t *>(qx_be + (size_t)i * C::BYTES);
    const uint32_t *y = reinterpret_cast<const uint32_t *>(qy_be + (size_t)i * C::BYTES);
    uint32_t h = seed;
#pragma unroll
    for (int k = 0; k < C::N; k += 2) {
        h = (h ^ __ldg(x + k)) * 0x9E3779B1u;
        h = (h ^ __ldg(y + k + 1)) * 0x85EBCA77u;
        h ^= h >> 15;
    }
    return h;
}
template <class C>
SBV_DEV bool kg_same_key(const uint8_t *__restrict__ qx_be, const uint8_t *__restrict__ qy_be, uint32_t i, uint32_t j) {
    const uint32_t *xi = reinterpret_cast<const uint32_t *>(qx_be + (size_t)i * C::BYTES);
    const uint32_t *xj = reinterpret_cast<const uint32_t *>(qx_be + (size_t)j * C::BYTES);
    const uint32_t *yi = reinterpret_cast<const uint32_t *>(qy_be + (size_t)i * C::BYTES);
    const uint32_t *yj = reinterpret_cast<const uint32_t *>(qy_be + (size_t)j * C::BYTES);
    uint32_t diff = 0;
#pragma unroll
    for (int k = 0; k < C::N; k++) diff |= (__ldg(xi + k) ^ __ldg(xj + k)) | (__ldg(yi + k) ^ __ldg(yj + k));
    return diff == 0;
}

// htab: hmask + 1 slots, all KG_EMPTY on entry; kcnt: n zeros on entry.
template <class C>
__global__ void __launch_bounds__(256) k_kg_insert(uint32_t n, const uint8_t *__restrict__ qx_be, const uint8_t *__restrict__ qy_be,
                                                   uint32_t seed, uint32_t hmask, uint32_t *__restrict__ htab,
                                                   uint32_t *__restrict__ rep, uint32_t *__restrict__ kcnt) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t h = kg_hash<C>(qx_be, qy_be, i, seed) & hmask;
    uint32_t r;
    for (;;) {
        uint32_t cur = htab[h];
        if (cur == KG_EMPTY) {
            cur = atomicCAS(htab + h, KG_EMPTY, i);
            if (cur == KG_EMPTY) { r = i; break; }
        }
        if (kg_same_key<C>(qx_be, qy_be, i, cur)) { r = cur; break; }
        h = (h + 1) & hmask;
    }
    rep[i] = r;
#if defined(__CUDA_ARCH__)
    // warp-aggregated count: a consensus batch has a handful of keys, i.e. thousands of items per counter
    const uint32_t peers = __match_any_sync(__activemask(), r);
    if ((threadIdx.x & 31) == (uint32_t)(__ffs((int)peers) - 1)) atomicAdd(kcnt + r, (uint32_t)__popc(peers));
#else
    atomicAdd(kcnt + r, 1u);
#endif
}

// keyid[i] (i a representative) = dense key id, or -1.  counters[0] = number of keys (may exceed max_keys: clamp
// when reading), counters[1] / counters[2] = fill of the fixed-base / generic lists; all zero on entry.
static __global__ void __launch_bounds__(256) k_kg_assign(uint32_t n, const uint32_t *__restrict__ rep, const uint32_t *__restrict__ kcnt,
                                                   uint32_t threshold, uint32_t max_keys, int32_t *__restrict__ keyid,
                                                   uint32_t *__restrict__ keylist, uint32_t *__restrict__ counters) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t id = -1;
    if (rep[i] == i && kcnt[i] >= threshold) {
        const uint32_t k = atomicAdd(counters + 0, 1u);
        if (k < max_keys) { id = (int32_t)k; keylist[k] = i; }
    }
    keyid[i] = id;
}

// item_kid[i] = key id of item i's key (or -1); klist / glist = the two work lists
static __global__ void __launch_bounds__(256) k_kg_route(uint32_t n, const uint32_t *__restrict__ rep, const int32_t *__restrict__ keyid,
                                                  int32_t *__restrict__ item_kid, uint32_t *__restrict__ klist, uint32_t *__restrict__ glist,
                                                  uint32_t *__restrict__ counters) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    const int32_t kid = live ? keyid[rep[i]] : -1;
    if (live) item_kid[i] = kid;
#if defined(__CUDA_ARCH__)
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t mk = __ballot_sync(0xffffffffu, live && kid >= 0), mg = __ballot_sync(0xffffffffu, live && kid < 0);
    uint32_t bk = 0, bg = 0;
    if (lane == 0) {
        if (mk) bk = atomicAdd(counters + 1, (uint32_t)__popc(mk));
        if (mg) bg = atomicAdd(counters + 2, (uint32_t)__popc(mg));
    }
    bk = __shfl_sync(0xffffffffu, bk, 0);
    bg = __shfl_sync(0xffffffffu, bg, 0);
    const uint32_t below = (1u << lane) - 1u;
    if (live && kid >= 0) klist[bk + __popc(mk & below)] = i;
    if (live && kid < 0) glist[bg + __popc(mg & below)] = i;
#else
    if (live && kid >= 0) klist[atomicAdd(counters + 1, 1u)] = i;
    if (live && kid < 0) glist[atomicAdd(counters + 2, 1u)] = i;
#endif
}

// ---- table construction -------------------------------------------------------------------------------------
// Scratch layout (cap = key capacity of the buffers; lanes of a warp are consecutive keys, so every access below
// is coalesced):
//   bases [win][3N words][cap]              Jacobian B_win = 2^(W*win) * Q
//   hs    [win][e = 2..ENT][N words][cap]   Z ratios of the co-Z chain: Z_e = Z_{e-1} * H_e  (H_2 = 2*Y_B, Z_1 = Z_B)
//   ztop  [win][N words][cap]               Z_ENT of the window; k_kt_inv overwrites it with its inverse
//   ktab  [kid][win][e-1][2N words]         Jacobian X, Y from k_kt_fill; affine x, y after k_kt_final

// nkeys_ptr: device counter (clamped to cap) — the grid is sized for the worst case and surplus threads leave.
// key k is item keylist[k] of (qx_be, qy_be); for registered keys keylist is the identity over the key array.
// INL: the eight multiplications of the doubling inlined (one site, ~25 KB): this kernel is a single dependent chain per
// thread on an otherwise idle SM sub-partition, so what counts is how well independent multiplications interleave.
template <class C, int W, bool INL>
__global__ void __launch_bounds__(64) k_kt_bases(const uint32_t *__restrict__ nkeys_ptr, uint32_t cap, const uint32_t *__restrict__ keylist,
                                                 const uint8_t *__restrict__ qx_be, const uint8_t *__restrict__ qy_be,
                                                 uint32_t *__restrict__ bases, uint8_t *__restrict__ keyflags) {
    constexpr int N = C::N;
    using KT = KeyTab<32 * N, W>;
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t nkeys = __ldg(nkeys_ptr);
    if (nkeys > cap) nkeys = cap;
    if (k >= nkeys) return;
    const uint32_t item = keylist ? keylist[k] : k;
    using A = typename PickArith<C, INL>::type;
    Jac<A> B;
    const bool good = load_key<C>(B.X, B.Y, qx_be, qy_be, item);
    keyflags[k] = good ? 1 : 0;
    if (!good) return;  // no table: every item of this key rejects (k_verify_kt checks the flag)
    C::get_one(B.Z);
#pragma unroll 1
    for (int win = 0; win < KT::NWIN; win++) {
        if (win) {
#pragma unroll 1
            for (int d = 0; d < W; d++) pt_double<A>(B);
        }
        uint32_t *o = bases + (size_t)win * 3 * N * cap + k;
#pragma unroll
        for (int i = 0; i < N; i++) { o[(size_t)i * cap] = B.X[i]; o[(size_t)(N + i) * cap] = B.Y[i]; o[(size_t)(2 * N + i) * cap] = B.Z[i]; }
    }
}

// k_kt_bases4 — the same chain with FOUR LANES PER KEY (three of them working): the eight multiplications of a doubling
// form four dependent levels, and the independent ones of a level run on different lanes:
//   level 1   lane 0: delta = Z*Z          lane 1: bb = (2Y)*(2Y)        lane 2: Z3 = (2Y)*Z
//   level 2   lane 0: (X-delta)*(X+delta)  lane 1: beta4 = X*bb          lane 2: bb*bb
//   level 3   lane 0: alpha*alpha          (alpha = 3*(X-delta)(X+delta))
//   level 4   lane 0: alpha*(beta4 - X3)
// with six 8-word quad broadcasts per doubling (bb, beta4, 8Y^4, and the new X, Y, Z).  The kernel is one dependent
// chain on an otherwise idle sub-partition, so halving the number of dependent multiplications halves its duration.
template <class C, int W>
__global__ void __launch_bounds__(128) k_kt_bases4(const uint32_t *__restrict__ nkeys_ptr, uint32_t cap, const uint32_t *__restrict__ keylist,
                                                   const uint8_t *__restrict__ qx_be, const uint8_t *__restrict__ qy_be,
                                                   uint32_t *__restrict__ bases, uint8_t *__restrict__ keyflags) {
    constexpr int N = C::N;
    using KT = KeyTab<32 * N, W>;
    const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t k = gt >> 2, role = gt & 3;
    uint32_t nkeys = __ldg(nkeys_ptr);
    if (nkeys > cap) nkeys = cap;
    if (k >= nkeys) return;  // a quad leaves together
    const unsigned qbase = (threadIdx.x & 31) & ~3u, qmask = 0xFu << qbase;
    const uint32_t item = keylist ? keylist[k] : k;
    uint32_t X[N], Y[N], Z[N];
    const bool good = load_key<C>(X, Y, qx_be, qy_be, item);  // the four lanes agree
    if (role == 0) keyflags[k] = good ? 1 : 0;
    if (!good) return;
    C::get_one(Z);
    auto bcast = [&](uint32_t (&v)[N], unsigned src) {
#pragma unroll
        for (int i = 0; i < N; i++) v[i] = __shfl_sync(qmask, v[i], qbase + src);
    };
#pragma unroll 1
    for (int win = 0; win < KT::NWIN; win++) {
        if (win) {
#pragma unroll 1
            for (int d = 0; d < W; d++) {
                uint32_t s[N], a[N], b[N], r1[N], r2[N], r3[N], r4[N], t1[N], t2[N];
                C::fadd(s, Y, Y);
                // level 1: delta | bb | Z3 | (delta)
                mp_select<N>(a, role == 1 || role == 2, s, Z);
                mp_select<N>(b, role == 1, s, Z);
                C::fmul(r1, a, b);
                uint32_t bb[N];
                mp_copy<N>(bb, r1);
                bcast(bb, 1);
                // level 2: (X-delta)(X+delta) | X*bb | bb*bb
                C::fsub(t1, X, r1);
                C::fadd(t2, X, r1);
                mp_select<N>(a, role == 0, t1, X);
                mp_select<N>(a, role == 2, bb, a);
                mp_select<N>(b, role == 0, t2, bb);
                C::fmul(r2, a, b);
                // level 3 (lane 0): alpha = 3*r2, alpha^2 ; lane 2: 8Y^4 = r2 / 2 ; lane 1 holds beta4 = r2
                uint32_t alpha[N], half[N], beta4[N];
                C::fadd(t1, r2, r2);
                C::fadd(alpha, t1, r2);
                C::fhalf(half, r2);
                C::fmul(r3, alpha, alpha);
                mp_copy<N>(beta4, r2);
                bcast(beta4, 1);
                bcast(half, 2);
                // level 4 (lane 0): X3 = alpha^2 - 2*beta4 ; Y3 = alpha*(beta4 - X3) - 8Y^4
                uint32_t x3[N], y3[N];
                C::fadd(t1, beta4, beta4);
                C::fsub(x3, r3, t1);
                C::fsub(t2, beta4, x3);
                C::fmul(r4, alpha, t2);
                C::fsub(y3, r4, half);
                mp_copy<N>(X, x3); bcast(X, 0);
                mp_copy<N>(Y, y3); bcast(Y, 0);
                mp_copy<N>(Z, r1); bcast(Z, 2);
            }
        }
        if (role == 0) {
            uint32_t *o = bases + (size_t)win * 3 * N * cap + k;
#pragma unroll
            for (int i = 0; i < N; i++) { o[(size_t)i * cap] = X[i]; o[(size_t)(N + i) * cap] = Y[i]; o[(size_t)(2 * N + i) * cap] = Z[i]; }
        }
    }
}

// co-Z addition (Meloni): P = (X1, Y1, Z) and Q = (X2, Y2, Z) share Z.  R = P + Q -> (X3, Y3, Z3) and P is
// re-expressed with the same Z3 = Z * (X1 - X2); h receives X1 - X2 (the ratio Z3 / Z).  5M + 2S.
// P != +-Q is the caller's business (multiples e*B, e >= 2, of a point of prime order never meet B).
template <class C>
SBV_DEV void zaddu(uint32_t (&X1)[C::N], uint32_t (&Y1)[C::N], const uint32_t (&X2)[C::N], const uint32_t (&Y2)[C::N],
                   uint32_t (&X3)[C::N], uint32_t (&Y3)[C::N], uint32_t (&h)[C::N]) {
    constexpr int N = C::N;
    uint32_t c[N], w1[N], w2[N], dy[N], d[N], a1[N], t[N];
    C::fsub(h, X1, X2);
    C::fsqr(c, h);
    C::fmul(w1, X1, c);
    C::fmul(w2, X2, c);
    C::fsub(dy, Y1, Y2);
    C::fsqr(d, dy);
    C::fsub(t, w1, w2);
    C::fmul(a1, Y1, t);
    C::fsub(X3, d, w1);
    C::fsub(X3, X3, w2);
    C::fsub(t, w1, X3);
    C::fmul(Y3, dy, t);
    C::fsub(Y3, Y3, a1);
    mp_copy<N>(X1, w1);
    mp_copy<N>(Y1, a1);
}

template <class C, int W>
__global__ void __launch_bounds__(64) k_kt_fill(const uint32_t *__restrict__ nkeys_ptr, uint32_t cap, const uint32_t *__restrict__ bases,
                                                const uint8_t *__restrict__ keyflags, uint32_t *__restrict__ hs,
                                                uint32_t *__restrict__ ztop, uint32_t *__restrict__ ktab) {
    constexpr int N = C::N;
    using KT = KeyTab<32 * N, W>;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t nkeys = __ldg(nkeys_ptr);
    if (nkeys > cap) nkeys = cap;
    if (t >= nkeys * KT::NWIN) return;
    const uint32_t k = t % nkeys, win = t / nkeys;
    if (!keyflags[k]) return;
    // B in Jacobian form with Z_B; entry 1 is B itself
    uint32_t bx[N], by[N], bz[N];
    {
        const uint32_t *o = bases + (size_t)win * 3 * N * cap + k;
#pragma unroll
        for (int i = 0; i < N; i++) { bx[i] = o[(size_t)i * cap]; by[i] = o[(size_t)(N + i) * cap]; bz[i] = o[(size_t)(2 * N + i) * cap]; }
    }
    uint32_t *out = ktab + ((size_t)k * KT::NWIN + win) * KT::ENT * 2 * N;
#pragma unroll
    for (int i = 0; i < N; i++) { out[i] = bx[i]; out[N + i] = by[i]; }
    // entry 2 = 2B (a = -3 Jacobian doubling); Z_2 = 2*Y_B*Z_B, so H_2 = 2*Y_B, and B is rescaled to Z_2:
    // (X_B * H^2, Y_B * H^3)
    Jac<C> P;
    mp_copy<N>(P.X, bx); mp_copy<N>(P.Y, by); mp_copy<N>(P.Z, bz);
    uint32_t h[N], h2[N], h3[N];
    C::fadd(h, by, by);
    pt_double<C>(P);
    C::fsqr(h2, h);
    C::fmul(h3, h2, h);
    C::fmul(bx, bx, h2);
    C::fmul(by, by, h3);
    uint32_t zacc[N];  // Z of the chain so far
    mp_copy<N>(zacc, P.Z);
    {
        uint32_t *hp = hs + ((size_t)win * (KT::ENT - 1) + 0) * N * cap + k;
#pragma unroll
        for (int i = 0; i < N; i++) { out[2 * N + i] = P.X[i]; out[3 * N + i] = P.Y[i]; hp[(size_t)i * cap] = h[i]; }
    }
    uint32_t px[N], py[N];
    mp_copy<N>(px, P.X); mp_copy<N>(py, P.Y);
#pragma unroll 1
    for (int e = 3; e <= KT::ENT; e++) {
        // (e)B = B + (e-1)B, both on the current Z; B is carried along to the new Z
        uint32_t x3[N], y3[N];
        zaddu<C>(bx, by, px, py, x3, y3, h);
        C::fmul(zacc, zacc, h);
        mp_copy<N>(px, x3); mp_copy<N>(py, y3);
        uint32_t *hp = hs + ((size_t)win * (KT::ENT - 1) + (e - 2)) * N * cap + k;
        uint32_t *oe = out + (size_t)(e - 1) * 2 * N;
#pragma unroll
        for (int i = 0; i < N; i++) { oe[i] = px[i]; oe[N + i] = py[i]; hp[(size_t)i * cap] = h[i]; }
    }
    uint32_t *zp = ztop + (size_t)win * N * cap + k;
#pragma unroll
    for (int i = 0; i < N; i++) zp[(size_t)i * cap] = zacc[i];
}

// ztop[win] <- 1 / ztop[win] for all windows of a key with one inversion
template <class C, int W>
__global__ void __launch_bounds__(64) k_kt_inv(const uint32_t *__restrict__ nkeys_ptr, uint32_t cap, const uint8_t *__restrict__ keyflags,
                                               uint32_t *__restrict__ ztop, uint32_t *__restrict__ pref) {
    constexpr int N = C::N;
    using KT = KeyTab<32 * N, W>;
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t nkeys = __ldg(nkeys_ptr);
    if (nkeys > cap) nkeys = cap;
    if (k >= nkeys || !keyflags[k]) return;
    uint32_t run[N];
    C::get_one(run);
#pragma unroll 1
    for (int win = 0; win < KT::NWIN; win++) {
        uint32_t z[N];
        const uint32_t *zp = ztop + (size_t)win * N * cap + k;
        uint32_t *pp = pref + (size_t)win * N * cap + k;
#pragma unroll
        for (int i = 0; i < N; i++) { z[i] = zp[(size_t)i * cap]; pp[(size_t)i * cap] = run[i]; }  // product of the windows before
        C::fmul(run, run, z);
    }
    uint32_t inv[N];
    p_inv<C>(inv, run);  // binary extended GCD: this thread is alone on its chain, the dependent length is what counts
#pragma unroll 1
    for (int win = KT::NWIN - 1; win >= 0; win--) {
        uint32_t z[N], pv[N], zi[N];
        uint32_t *zp = ztop + (size_t)win * N * cap + k;
        const uint32_t *pp = pref + (size_t)win * N * cap + k;
#pragma unroll
        for (int i = 0; i < N; i++) { z[i] = zp[(size_t)i * cap]; pv[i] = pp[(size_t)i * cap]; }
        C::fmul(zi, inv, pv);
        C::fmul(inv, inv, z);
#pragma unroll
        for (int i = 0; i < N; i++) zp[(size_t)i * cap] = zi[i];
    }
}

template <class C, int W>
__global__ void __launch_bounds__(64) k_kt_final(const uint32_t *__restrict__ nkeys_ptr, uint32_t cap, const uint32_t *__restrict__ bases,
                                                 const uint8_t *__restrict__ keyflags, const uint32_t *__restrict__ hs,
                                                 const uint32_t *__restrict__ ztop, uint32_t *__restrict__ ktab) {
    constexpr int N = C::N;
    using KT = KeyTab<32 * N, W>;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t nkeys = __ldg(nkeys_ptr);
    if (nkeys > cap) nkeys = cap;
    if (t >= nkeys * KT::NWIN) return;
    const uint32_t k = t % nkeys, win = t / nkeys;
    if (!keyflags[k]) return;
    uint32_t zi[N];  // 1 / Z_e, walking e = ENT .. 1
    {
        const uint32_t *zp = ztop + (size_t)win * N * cap + k;
#pragma unroll
        for (int i = 0; i < N; i++) zi[i] = zp[(size_t)i * cap];
    }
    uint32_t *out = ktab + ((size_t)k * KT::NWIN + win) * KT::ENT * 2 * N;
    // entry e-1 and its Z ratio are loaded before entry e is converted and stored (independent addresses: the loads
    // overlap the six multiplications)
    uint32_t x[N], y[N], h[N];
    {
        const uint32_t *oe = out + (size_t)(KT::ENT - 1) * 2 * N;
#pragma unroll
        for (int i = 0; i < N; i++) { x[i] = oe[i]; y[i] = oe[N + i]; }
    }
#pragma unroll 1
    for (int e = KT::ENT; e >= 1; e--) {
        uint32_t nx[N], ny[N], nh[N];
        if (e >= 2) {
            const uint32_t *on = out + (size_t)(e - 2) * 2 * N;
            const uint32_t *hp = hs + ((size_t)win * (KT::ENT - 1) + (e - 2)) * N * cap + k;
#pragma unroll
            for (int i = 0; i < N; i++) { nx[i] = on[i]; ny[i] = on[N + i]; nh[i] = hp[(size_t)i * cap]; }
        }
        uint32_t *oe = out + (size_t)(e - 1) * 2 * N;
        uint32_t z2[N], z3[N];
        C::fsqr(z2, zi);
        C::fmul(z3, z2, zi);
        C::fmul(x, x, z2);
        C::fmul(y, y, z3);
#pragma unroll
        for (int i = 0; i < N; i++) { oe[i] = x[i]; oe[N + i] = y[i]; }
        if (e >= 2) {  // 1/Z_{e-1} = (1/Z_e) * H_e
            mp_copy<N>(h, nh);
            C::fmul(zi, zi, h);
            mp_copy<N>(x, nx); mp_copy<N>(y, ny);
        }
    }
    (void)bases;
}

// words of scratch the builder needs for `cap` keys
template <class C, int W>
struct KtSizes {
    using KT = KeyTab<32 * C::N, W>;
    static constexpr size_t bases_words(size_t cap) { return (size_t)KT::NWIN * 3 * C::N * cap; }
    static constexpr size_t hs_words(size_t cap) { return (size_t)KT::NWIN * (KT::ENT - 1) * C::N * cap; }
    static constexpr size_t ztop_words(size_t cap) { return (size_t)KT::NWIN * C::N * cap; }
    static constexpr size_t ktab_words(size_t cap) { return KT::POINTS * 2 * C::N * cap; }
};

}  // namespace sbv
