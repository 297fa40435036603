// hostsim.cpp — TEST INFRASTRUCTURE: the device headers of libsbv compiled for the CPU (see csrc/hostsim.h) and
// driven one simulated thread at a time.  tests/test_hostsim.py (-m "not gpu") compares the results with Python
// big integers and with the oracle, so limb-level mistakes are caught without a GPU.  Not linked into libsbv.so.
#define SBV_HOSTSIM 1
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <thread>
#include <vector>

#include "../../consensus_b200/csrc/hostsim.h"
thread_local hostsim_dim3 threadIdx, blockIdx, blockDim, gridDim;
thread_local hostsim_warp *hostsim_ctx = nullptr;
namespace sbv { uint32_t tab[1 << 18]; }
#include "../../consensus_b200/csrc/debug_ops.cuh"
#include "../../consensus_b200/csrc/keygroup.cuh"
#include "../../consensus_b200/csrc/sha256.cuh"
#include "../../consensus_b200/csrc/quorum.cuh"

using namespace sbv;

// G comb for the simulation: T[i][b] = b * 2^(GW*i) * G, built incrementally (running sum + one batched inversion per
// window) — the device kernel k_gtable_init builds every entry independently, which a CPU cannot afford.
template <class C>
static void build_comb_host(uint32_t *tab) {
    constexpr int N = C::N;
    const size_t per = (size_t)1 << C::GW;
    Jac<C> base;
    C::get_gx(base.X); C::get_gy(base.Y); C::get_one(base.Z);
    std::vector<uint32_t> jac(per * 3 * N), pref(per * N);
    for (int win = 0; win < C::GWINS; win++) {
        if (win) for (int d = 0; d < C::GW; d++) pt_double<C>(base);
        Jac<C> acc = base;
        uint32_t run[N];
        C::get_one(run);
        for (size_t b = 1; b < per; b++) {
            if (b > 1) pt_add<C, false>(acc, base.X, base.Y, base.Z, false, false);
            memcpy(&jac[b * 3 * N], acc.X, 4 * N); memcpy(&jac[b * 3 * N + N], acc.Y, 4 * N); memcpy(&jac[b * 3 * N + 2 * N], acc.Z, 4 * N);
            memcpy(&pref[b * N], run, 4 * N);
            C::fmul(run, run, acc.Z);
        }
        uint32_t inv[N];
        f_inv<C>(inv, run);
        uint32_t *out = tab + ((size_t)win << C::GW) * 2 * N;
        memset(out, 0, 2 * N * 4);
        for (size_t b = per - 1; b >= 1; b--) {
            uint32_t pv[N], z[N], zi[N], z2[N], z3[N], x[N], y[N];
            memcpy(pv, &pref[b * N], 4 * N); memcpy(z, &jac[b * 3 * N + 2 * N], 4 * N);
            memcpy(x, &jac[b * 3 * N], 4 * N); memcpy(y, &jac[b * 3 * N + N], 4 * N);
            C::fmul(zi, inv, pv);
            C::fmul(inv, inv, z);
            C::fsqr(z2, zi); C::fmul(z3, z2, zi);
            C::fmul(x, x, z2); C::fmul(y, y, z3);
            memcpy(out + b * 2 * N, x, 4 * N); memcpy(out + b * 2 * N + N, y, 4 * N);
        }
    }
}

template <class F>
static void run_grid(unsigned blocks, unsigned threads, F &&body) {
    gridDim.x = blocks; blockDim.x = threads;
    for (unsigned b = 0; b < blocks; b++)
        for (unsigned t = 0; t < threads; t++) { blockIdx.x = b; threadIdx.x = t; body(); }
}

extern "C" int hs_debug_op(int curve, int op, size_t n, const uint32_t *a, const uint32_t *b, uint32_t *out) {
    for (size_t i = 0; i < n; i++) {
        if (curve == 0) debug_op_item<P256>(op, (uint32_t)i, a, b, out);
        else debug_op_item<P384>(op, (uint32_t)i, a, b, out);
    }
    return 0;
}

// LOCKSTEP form for warp-cooperative kernels: the 32 lanes of a warp are 32 OS threads that meet at every shuffle / ballot
// (hostsim.h); warps run one after the other.
template <class F>
static void run_grid_lockstep(unsigned blocks, unsigned threads, F &&body) {
    for (unsigned b = 0; b < blocks; b++)
        for (unsigned w0 = 0; w0 < threads; w0 += 32) {
            hostsim_warp ctx;
            std::vector<std::thread> lanes;
            for (unsigned t = w0; t < w0 + 32 && t < threads; t++)
                lanes.emplace_back([&, t] {
                    gridDim.x = blocks; blockDim.x = threads; blockIdx.x = b; threadIdx.x = t;
                    hostsim_ctx = &ctx;
                    body();
                    hostsim_ctx = nullptr;
                });
            for (auto &l : lanes) l.join();
        }
}

// per-key tables of `nkeys` keys (key k = item k of qx / qy), built by the product's four kernels; four != 0: the doubling
// chain by k_kt_bases4 (four lanes per key, in lockstep), else by the one-thread-per-key k_kt_bases.  ktab_out: the final
// affine tables (KtSizes::ktab_words(nkeys) words), flags_out: nkeys validity flags.
template <class C, int W>
static void tables_t(uint32_t nkeys, const uint8_t *qx, const uint8_t *qy, int four, uint32_t *ktab_out, uint8_t *flags_out) {
    using KS = KtSizes<C, W>;
    using KT = KeyTab<32 * C::N, W>;
    const size_t cap = nkeys;
    std::vector<uint32_t> bases(KS::bases_words(cap)), hs(KS::hs_words(cap)), ztop(KS::ztop_words(cap)), pref(KS::ztop_words(cap)), ktab(KS::ktab_words(cap), 0);
    std::vector<uint8_t> kflags(cap, 0);
    uint32_t cnt = nkeys;
    if (four) run_grid_lockstep((unsigned)((cap * 4 + 127) / 128), 128, [&] { k_kt_bases4<C, W>(&cnt, (uint32_t)cap, nullptr, qx, qy, bases.data(), kflags.data()); });
    else run_grid((unsigned)((cap + 63) / 64), 64, [&] { k_kt_bases<C, W, true>(&cnt, (uint32_t)cap, nullptr, qx, qy, bases.data(), kflags.data()); });
    run_grid((unsigned)((cap * KT::NWIN + 63) / 64), 64, [&] { k_kt_fill<C, W>(&cnt, (uint32_t)cap, bases.data(), kflags.data(), hs.data(), ztop.data(), ktab.data()); });
    run_grid((unsigned)((cap + 63) / 64), 64, [&] { k_kt_inv<C, W>(&cnt, (uint32_t)cap, kflags.data(), ztop.data(), pref.data()); });
    run_grid((unsigned)((cap * KT::NWIN + 63) / 64), 64, [&] { k_kt_final<C, W>(&cnt, (uint32_t)cap, bases.data(), kflags.data(), hs.data(), ztop.data(), ktab.data()); });
    memcpy(ktab_out, ktab.data(), ktab.size() * 4);
    memcpy(flags_out, kflags.data(), cap);
}

extern "C" size_t hs_ktab_words(int curve, int w8, size_t nkeys) {
    if (curve == 0) return w8 ? KtSizes<P256, 8>::ktab_words(nkeys) : KtSizes<P256, 5>::ktab_words(nkeys);
    return w8 ? KtSizes<P384, 8>::ktab_words(nkeys) : KtSizes<P384, 5>::ktab_words(nkeys);
}
extern "C" int hs_tables(int curve, int w8, size_t nkeys, const uint8_t *qx, const uint8_t *qy, int four, uint32_t *ktab_out, uint8_t *flags_out) {
    if (curve == 0 && !w8) tables_t<P256, 5>((uint32_t)nkeys, qx, qy, four, ktab_out, flags_out);
    else if (curve == 0) tables_t<P256, 8>((uint32_t)nkeys, qx, qy, four, ktab_out, flags_out);
    else if (!w8) tables_t<P384, 5>((uint32_t)nkeys, qx, qy, four, ktab_out, flags_out);
    else tables_t<P384, 8>((uint32_t)nkeys, qx, qy, four, ktab_out, flags_out);
    return 0;
}

// k_sha256 over a ragged batch, one message per simulated thread (perm: optional processing order, as the counting sort gives it)
extern "C" int hs_sha256(size_t n, const uint8_t *msgs, const uint64_t *off, uint64_t base, const uint32_t *perm, uint8_t *digest_out) {
    run_grid((unsigned)((n + 127) / 128), 128, [&] { k_sha256((uint32_t)n, msgs, off, base, digest_out, perm); });
    return 0;
}

// quorum counting (k_quorum_count + k_quorum_reached) for the instances [inst_base, inst_base + n_instances), as one device of a
// sharded engine runs it; ok may be NULL (prepares)
extern "C" int hs_quorum(size_t n_votes, const uint32_t *instance, const uint16_t *sender, const uint16_t *signer, const uint8_t *digest_match,
                         const uint8_t *ok, const uint16_t *self_id, uint32_t inst_base, size_t n_instances, uint32_t threshold, uint32_t *valid_count,
                         uint8_t *reached) {
    memset(valid_count, 0, n_instances * 4);
    run_grid((unsigned)((n_votes + 255) / 256), 256, [&] {
        k_quorum_count((uint32_t)n_votes, instance, sender, signer, digest_match, ok, self_id, inst_base, (uint32_t)n_instances, valid_count);
    });
    run_grid((unsigned)((n_instances + 255) / 256), 256, [&] { k_quorum_reached((uint32_t)n_instances, valid_count, threshold, reached); });
    return 0;
}
// k_pack_bits in lockstep (a warp ballot per 32 verdicts)
extern "C" int hs_pack_bits(size_t n, const uint8_t *ok, uint32_t *mask) {
    run_grid_lockstep((unsigned)((n + 255) / 256), 256, [&] { k_pack_bits((uint32_t)n, ok, mask); });
    return 0;
}

// registered-key path (sbv_set_keys / sbv_verify_registered): 8-bit window tables for `nkeys` keys, then k_prep and the
// fixed-base kernel with the key taken by slot — thread per signature, or (warp != 0) ONE SIGNATURE PER WARP in lockstep
// (k_verify_kt_warp: lanes add their table points, shuffle-tree reduction)
template <class C>
static void registered_t(uint32_t n, uint32_t nkeys, const uint8_t *kx, const uint8_t *ky, const uint32_t *slot, const uint8_t *r, const uint8_t *s,
                         const uint8_t *dig, uint32_t dlen, const uint4 *gtab, int warp, uint8_t *ok) {
    constexpr int N = C::N, S = 8, W = 8;
    using KS = KtSizes<C, W>;
    std::vector<uint32_t> ktab(KS::ktab_words(nkeys));
    std::vector<uint8_t> kflags(nkeys);
    tables_t<C, W>(nkeys, kx, ky, 0, ktab.data(), kflags.data());
    std::vector<int32_t> s2l(nkeys);
    for (uint32_t i = 0; i < nkeys; i++) s2l[i] = (int32_t)i;
    std::vector<uint32_t> uw((size_t)2 * N * n);
    std::vector<uint8_t> flags(n);
    run_grid(((n + S - 1) / S + 127) / 128, 128, [&] { k_prep<C, S>(n, r, s, dig, dlen, uw.data(), flags.data()); });
    const uint4 *k4 = reinterpret_cast<const uint4 *>(ktab.data());
    if (warp)
        run_grid_lockstep((unsigned)(((size_t)n * 32 + 127) / 128), 128,
                          [&] { k_verify_kt_warp<C, W>(n, slot, s2l.data(), nkeys, kflags.data(), r, uw.data(), flags.data(), gtab, k4, ok); });
    else
        run_grid((n + 63) / 64, 64, [&] {
            k_verify_kt<C, W, 64, 1, true, false>(n, slot, s2l.data(), nkeys, kflags.data(), r, uw.data(), flags.data(), gtab, k4, ok, nullptr, nullptr, nullptr);
        });
}

template <class C> static const uint4 *gtab_for(int idx);
extern "C" int hs_verify_registered(int curve, size_t n, size_t nkeys, const uint8_t *kx, const uint8_t *ky, const uint32_t *slot, const uint8_t *r,
                                    const uint8_t *s, const uint8_t *dig, uint32_t dlen, int warp, uint8_t *ok) {
    if (curve == 0) registered_t<P256>((uint32_t)n, (uint32_t)nkeys, kx, ky, slot, r, s, dig, dlen, gtab_for<P256>(0), warp, ok);
    else registered_t<P384>((uint32_t)n, (uint32_t)nkeys, kx, ky, slot, r, s, dig, dlen, gtab_for<P384>(1), warp, ok);
    return 0;
}

// keys-per-item path: k_prep + k_verify_coz, exactly the kernels of the product, thread by thread
template <class C, int BLOCK>
static void verify_coz_t(uint32_t n, const uint8_t *r, const uint8_t *s, const uint8_t *qx, const uint8_t *qy, const uint8_t *dig,
                         uint32_t dlen, const uint4 *gtab, uint8_t *ok) {
    constexpr int N = C::N, S = 8;
    std::vector<uint32_t> uw((size_t)2 * N * n);
    std::vector<uint8_t> flags(n);
    std::vector<uint32_t> tscr((size_t)12 * N * n);
    const unsigned pthreads = (n + S - 1) / S;
    run_grid((pthreads + 127) / 128, 128, [&] { k_prep<C, S>(n, r, s, dig, dlen, uw.data(), flags.data()); });
    run_grid((n + BLOCK - 1) / BLOCK, BLOCK, [&] {
        k_verify_coz<C, BLOCK, 1>(n, qx, qy, r, uw.data(), flags.data(), gtab, tscr.data(), ok, nullptr, nullptr);
    });
}

// grouped path: k_prep, key grouping, table construction, k_verify_kt for repeated keys + k_verify_coz for the rest
template <class C, int W>
static void verify_grouped_t(uint32_t n, const uint8_t *r, const uint8_t *s, const uint8_t *qx, const uint8_t *qy, const uint8_t *dig,
                             uint32_t dlen, const uint4 *gtab, uint32_t threshold, uint32_t max_keys, uint8_t *ok, uint32_t *stats,
                             uint32_t chunk = 0) {
    constexpr int N = C::N, S = 8;
    using KS = KtSizes<C, W>;
    using KT = KeyTab<32 * N, W>;
    std::vector<uint32_t> uw((size_t)2 * N * n);
    std::vector<uint8_t> flags(n);
    std::vector<uint32_t> tscr((size_t)12 * N * n);
    if (!chunk) run_grid(((n + S - 1) / S + 127) / 128, 128, [&] { k_prep<C, S>(n, r, s, dig, dlen, uw.data(), flags.data()); });
    uint32_t hsize = 1;
    while (hsize < 2 * n) hsize <<= 1;
    std::vector<uint32_t> htab(hsize, KG_EMPTY), rep(n), kcnt(n, 0), keylist(max_keys ? max_keys : 1), klist(n), glist(n), counters(4, 0);
    std::vector<int32_t> keyid(n), item_kid(n);
    run_grid((n + 255) / 256, 256, [&] { k_kg_insert<C>(n, qx, qy, 0x1234567u, hsize - 1, htab.data(), rep.data(), kcnt.data()); });
    run_grid((n + 255) / 256, 256, [&] { k_kg_assign(n, rep.data(), kcnt.data(), threshold, max_keys, keyid.data(), keylist.data(), counters.data()); });
    if (!chunk) run_grid((n + 255) / 256, 256, [&] { k_kg_route(n, rep.data(), keyid.data(), item_kid.data(), klist.data(), glist.data(), counters.data()); });
    const size_t cap = max_keys ? max_keys : 1;
    std::vector<uint32_t> bases(KS::bases_words(cap)), hs(KS::hs_words(cap)), ztop(KS::ztop_words(cap)), pref(KS::ztop_words(cap)), ktab(KS::ktab_words(cap));
    std::vector<uint8_t> kflags(cap, 0);
    run_grid((unsigned)((cap + 63) / 64), 64, [&] { k_kt_bases<C, W, true>(counters.data(), (uint32_t)cap, keylist.data(), qx, qy, bases.data(), kflags.data()); });
    run_grid((unsigned)((cap * KT::NWIN + 63) / 64), 64, [&] { k_kt_fill<C, W>(counters.data(), (uint32_t)cap, bases.data(), kflags.data(), hs.data(), ztop.data(), ktab.data()); });
    run_grid((unsigned)((cap + 63) / 64), 64, [&] { k_kt_inv<C, W>(counters.data(), (uint32_t)cap, kflags.data(), ztop.data(), pref.data()); });
    run_grid((unsigned)((cap * KT::NWIN + 63) / 64), 64, [&] { k_kt_final<C, W>(counters.data(), (uint32_t)cap, bases.data(), kflags.data(), hs.data(), ztop.data(), ktab.data()); });
    const bool gsplit = (threshold & 1) == 0;  // exercise both forms: even thresholds take the split u1*G path
    std::vector<uint32_t> gacc((size_t)3 * N * n);
    if (chunk) {
        // The second half chunk by chunk, as sbv_launch_verify_chunk (csrc/pipeline.cu) enqueues it: the items [lo, lo + cn) are a
        // batch of their own for every per-item array (word-major, stride = the chunk's size, base = words per item * lo); the
        // grouping (rep, keyid) and the key tables are shared; routing is per chunk, chunk-local indices, the chunk's own counters.
        const size_t L = C::BYTES;
        uint32_t kt_total = 0, gen_total = 0;
        for (uint32_t lo = 0; lo < n; lo += chunk) {
            const uint32_t cn = n - lo < chunk ? n - lo : chunk;
            std::vector<uint32_t> cc(4, 0);
            uint32_t *uwc = uw.data() + (size_t)2 * N * lo, *tsc = tscr.data() + (size_t)12 * N * lo, *gac = gacc.data() + (size_t)3 * N * lo;
            uint8_t *flc = flags.data() + lo;
            const uint8_t *rc = r + lo * L;
            run_grid(((cn + S - 1) / S + 127) / 128, 128, [&] { k_prep<C, S>(cn, rc, s + lo * L, dig + (size_t)lo * dlen, dlen, uwc, flc); });
            run_grid((cn + 255) / 256, 256, [&] { k_kg_route(cn, rep.data() + lo, keyid.data(), item_kid.data() + lo, klist.data() + lo, glist.data() + lo, cc.data()); });
            run_grid((cn + 63) / 64, 64, [&] { k_verify_coz<C, 64, 1>(cn, qx + lo * L, qy + lo * L, rc, uwc, flc, gtab, tsc, ok + lo, glist.data() + lo, cc.data() + 2); });
            if (gsplit) run_grid((cn + 63) / 64, 64, [&] { k_gpart<C, 64, 1>(cn, uwc, gtab, gac); });
            run_grid((cn + 63) / 64, 64, [&] {
                k_verify_kt<C, W, 64, 1, false, false>(cn, nullptr, item_kid.data() + lo, 0, kflags.data(), rc, uwc, flc, gtab,
                                                reinterpret_cast<const uint4 *>(ktab.data()), ok + lo, klist.data() + lo, cc.data() + 1, gsplit ? gac : nullptr);
            });
            kt_total += cc[1]; gen_total += cc[2];
        }
        if (stats) { stats[0] = counters[0]; stats[1] = kt_total; stats[2] = gen_total; }
        return;
    }
    if (gsplit) run_grid((n + 63) / 64, 64, [&] { k_gpart<C, 64, 1>(n, uw.data(), gtab, gacc.data()); });
    run_grid((n + 63) / 64, 64, [&] {
        k_verify_kt<C, W, 64, 1, false, false>(n, nullptr, item_kid.data(), 0, kflags.data(), r, uw.data(), flags.data(), gtab,
                                        reinterpret_cast<const uint4 *>(ktab.data()), ok, klist.data(), counters.data() + 1, gsplit ? gacc.data() : nullptr);
    });
    run_grid((n + 63) / 64, 64, [&] {
        k_verify_coz<C, 64, 1>(n, qx, qy, r, uw.data(), flags.data(), gtab, tscr.data(), ok, glist.data(), counters.data() + 2);
    });
    if (stats) { stats[0] = counters[0]; stats[1] = counters[1]; stats[2] = counters[2]; }
}

static std::vector<uint32_t> g_gtab[2];
template <class C>
static const uint4 *gtab_for(int idx) {
    auto &g = g_gtab[idx];
    if (g.empty()) {
        const size_t entries = (size_t)C::GWINS << C::GW;
        g.resize(entries * 2 * C::N);
        // incremental construction (the device kernel builds every entry independently, far too slow for a CPU)
        build_comb_host<C>(g.data());
    }
    return reinterpret_cast<const uint4 *>(g.data());
}

// the same with the second half run chunk by chunk (chunk = items per chunk), as a chunked host-buffer call does
extern "C" int hs_verify_chunked(int curve, size_t n, const uint8_t *r, const uint8_t *s, const uint8_t *qx, const uint8_t *qy, const uint8_t *dig,
                      uint32_t dlen, uint32_t threshold, uint32_t max_keys, uint32_t chunk, uint8_t *ok, uint32_t *stats) {
    if (curve == 0) verify_grouped_t<P256, 5>((uint32_t)n, r, s, qx, qy, dig, dlen, gtab_for<P256>(0), threshold, max_keys, ok, stats, chunk);
    else verify_grouped_t<P384, 5>((uint32_t)n, r, s, qx, qy, dig, dlen, gtab_for<P384>(1), threshold, max_keys, ok, stats, chunk);
    return 0;
}

extern "C" int hs_verify(int curve, size_t n, const uint8_t *r, const uint8_t *s, const uint8_t *qx, const uint8_t *qy, const uint8_t *dig, uint32_t dlen,
              uint8_t *ok) {
    if (curve == 0) verify_coz_t<P256, 64>((uint32_t)n, r, s, qx, qy, dig, dlen, gtab_for<P256>(0), ok);
    else verify_coz_t<P384, 64>((uint32_t)n, r, s, qx, qy, dig, dlen, gtab_for<P384>(1), ok);
    return 0;
}

// grouped (fixed-base for repeated keys) path; stats = {keys found, items on the fixed-base path, items on the generic path}
extern "C" int hs_verify_grouped(int curve, size_t n, const uint8_t *r, const uint8_t *s, const uint8_t *qx, const uint8_t *qy, const uint8_t *dig,
                      uint32_t dlen, uint32_t threshold, uint32_t max_keys, uint8_t *ok, uint32_t *stats) {
    if (curve == 0) verify_grouped_t<P256, 5>((uint32_t)n, r, s, qx, qy, dig, dlen, gtab_for<P256>(0), threshold, max_keys, ok, stats);
    else verify_grouped_t<P384, 5>((uint32_t)n, r, s, qx, qy, dig, dlen, gtab_for<P384>(1), threshold, max_keys, ok, stats);
    return 0;
}

