// pipeline.cu — the verify pipelines of libsbv.so: which kernels run for a batch, on which streams, over which scratch.
//
// Keys-per-item batch (sbv_verify_batch*, sbv_hash_verify_batch, sbv_verify_mixed):
//
//   st     memsets  k_kg_insert  k_kg_assign  k_kg_route ─┬─ k_prep ───────────────┬─ (wait tables) k_verify_kt ─ (wait generic) ─ done
//   s_tab                                                 └─ k_kt_bases  k_kt_fill  k_kt_inv  k_kt_final ─┘
//   s_gen                                                                          └─ k_verify_coz (keys without a table) ─┘
//
// Keys that occur at least `group_threshold` times in the batch get a fixed-base table built on the spot (keygroup.cuh)
// and their signatures take the doubling-free kernel; the rest take the generic kernel.  The scalar preparation
// (latency-bound: one inversion chain) runs beside the table construction (latency-bound: one doubling chain).
// Registered keys (sbv_set_keys) skip the grouping: their tables were built at registration.
#include "engine.h"

namespace {

// Takes the next scratch set of device d for a launch on stream st: waits (on the stream) for the set's previous user
// and grows the buffers to n items / kcap keys of curve `ops` (growth drains the previous user on the host first).
int take_scratch(sbv_engine *e, Dev &d, const CurveOps &ops, const KtOps *kt, size_t n, size_t kcap, cudaStream_t st, Dev::Scratch **out) {
    // round robin over the sets that are not held open between the two halves of a host-buffer launch (at most
    // SBV_LANES < SBV_SCRATCH of them at any time)
    int idx = (int)(d.ws_next++ % SBV_SCRATCH);
    for (int tries = 0; tries < SBV_SCRATCH && d.ws[idx].open; tries++) idx = (int)(d.ws_next++ % SBV_SCRATCH);
    if (d.ws[idx].open) return sbv_fail(e, SBV_ERR_ARG, "no free scratch set (more launches held open than lanes?)");
    Dev::Scratch &w = d.ws[idx];
    Dev::Scratch::Caps &c = w.caps;
    if (!w.done) {
        // The events and side streams of EVERY set are created now: a set first taken in the middle of a steady stream of
        // launches would otherwise stop to create streams there (measured: 2 - 40 ms at the head of a timed region).
        for (int j = 0; j < SBV_SCRATCH; j++) {
            Dev::Scratch &x = d.ws[j];
            if (x.done) continue;
            CU(e, cudaEventCreateWithFlags(&x.done, cudaEventDisableTiming));
            CU(e, cudaEventCreateWithFlags(&x.ev_group, cudaEventDisableTiming));
            CU(e, cudaEventCreateWithFlags(&x.ev_prep, cudaEventDisableTiming));
            CU(e, cudaEventCreateWithFlags(&x.ev_tab, cudaEventDisableTiming));
            CU(e, cudaEventCreateWithFlags(&x.ev_gen, cudaEventDisableTiming));
            if (e->tab_hi) {
                int lo_p = 0, hi_p = 0;
                CU(e, cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));
                CU(e, cudaStreamCreateWithPriority(&x.s_tab, cudaStreamNonBlocking, hi_p));
            } else {
                CU(e, cudaStreamCreateWithFlags(&x.s_tab, cudaStreamNonBlocking));
            }
            CU(e, cudaStreamCreateWithFlags(&x.s_gen, cudaStreamNonBlocking));
        }
    }
    const size_t N = (size_t)ops.N;
    uint32_t hsize = 1;
    while (hsize < 2 * n) hsize <<= 1;
    struct Need { void **p; size_t *cap; size_t need, alloc; };
    // `need` = bytes this launch uses; `alloc` = bytes to allocate when the buffer must grow (headroom so that a slowly
    // growing batch size does not reallocate every call)
    const size_t ni = n + n / 8 + 1024, kc = kcap + kcap / 8 + 16;
    uint32_t hs2 = 1;
    while (hs2 < 2 * ni) hs2 <<= 1;
    const bool g = kcap > 0, t = g && kt;
    const KtGeom z{};
    const KtGeom &q = t ? kt->geom : z;
    Need needs[] = {
        {(void **)&w.uw, &c.uw, 2 * N * n * 4, 2 * N * ni * 4},
        {(void **)&w.flags, &c.flags, n, ni},
        {(void **)&w.tscr, &c.tscr, 12 * N * n * 4, 12 * N * ni * 4},
        {(void **)&w.gacc, &c.gacc, g ? 3 * N * n * 4 : 0, 3 * N * ni * 4},
        {(void **)&w.htab, &c.htab, g ? (size_t)hsize * 4 : 0, (size_t)hs2 * 4},
        {(void **)&w.rep, &c.rep, g ? n * 4 : 0, ni * 4},
        {(void **)&w.klist, &c.klist, g ? n * 4 : 0, ni * 4},
        {(void **)&w.glist, &c.glist, g ? n * 4 : 0, ni * 4},
        {(void **)&w.zeroed, &c.zeroed, g ? (n + 4 + 4 * SBV_MAX_CHUNKS) * 4 : 0, (ni + 4 + 4 * SBV_MAX_CHUNKS) * 4},
        {(void **)&w.keyid, &c.keyid, g ? n * 4 : 0, ni * 4},
        {(void **)&w.item_kid, &c.item_kid, g ? n * 4 : 0, ni * 4},
        {(void **)&w.keylist, &c.keylist, t ? kcap * 4 : 0, kc * 4},
        {(void **)&w.keyflags, &c.keyflags, t ? kcap : 0, kc},
        {(void **)&w.bases, &c.bases, q.bases_words * kcap * 4, q.bases_words * kc * 4},
        {(void **)&w.hs, &c.hs, q.hs_words * kcap * 4, q.hs_words * kc * 4},
        {(void **)&w.ztop, &c.ztop, q.ztop_words * kcap * 4, q.ztop_words * kc * 4},
        {(void **)&w.pref, &c.pref, q.ztop_words * kcap * 4, q.ztop_words * kc * 4},
        {(void **)&w.ktab, &c.ktab, q.ktab_words * kcap * 4, q.ktab_words * kc * 4},
    };
    bool grows = false;
    for (const Need &nd : needs) grows = grows || nd.need > *nd.cap;
    if (grows) {
        // Grow EVERY scratch set of the device now, not just this one: otherwise the first launch on each of the other sets
        // would stop to allocate in the middle of a steady stream of launches (cudaMalloc of hundreds of MB synchronises).
        for (int j = 0; j < SBV_SCRATCH; j++) {
            Dev::Scratch &x = d.ws[j];
            if (x.open && &x != &w) continue;  // held by a launch between its halves: it grows when it is next taken
            if (x.used && x.done) CU(e, cudaEventSynchronize(x.done));  // nothing may still be using the buffers we are about to free
            Dev::Scratch::Caps &cx = x.caps;
            struct Slot { void **p; size_t *cap; };
            const Slot slots[] = {
                {(void **)&x.uw, &cx.uw}, {(void **)&x.flags, &cx.flags}, {(void **)&x.tscr, &cx.tscr}, {(void **)&x.gacc, &cx.gacc},
                {(void **)&x.htab, &cx.htab}, {(void **)&x.rep, &cx.rep}, {(void **)&x.klist, &cx.klist}, {(void **)&x.glist, &cx.glist},
                {(void **)&x.zeroed, &cx.zeroed}, {(void **)&x.keyid, &cx.keyid}, {(void **)&x.item_kid, &cx.item_kid}, {(void **)&x.keylist, &cx.keylist},
                {(void **)&x.keyflags, &cx.keyflags}, {(void **)&x.bases, &cx.bases}, {(void **)&x.hs, &cx.hs}, {(void **)&x.ztop, &cx.ztop},
                {(void **)&x.pref, &cx.pref}, {(void **)&x.ktab, &cx.ktab},
            };
            static_assert(sizeof(slots) / sizeof(slots[0]) == sizeof(needs) / sizeof(needs[0]), "one slot per buffer");
            for (size_t q = 0; q < sizeof(slots) / sizeof(slots[0]); q++) {
                if (needs[q].need <= *slots[q].cap) continue;
                if (*slots[q].p) cudaFree(*slots[q].p);
                *slots[q].p = nullptr;
                *slots[q].cap = 0;
                CU(e, cudaMalloc(slots[q].p, needs[q].alloc));
                *slots[q].cap = needs[q].alloc;
            }
        }
    }
    w.hsize = hsize;
    if (w.used) CU(e, cudaStreamWaitEvent(st, w.done, 0));
    w.used = true;
    *out = &w;
    return 0;
}

cudaEvent_t *prof_take(sbv_engine *e, Dev &d) {
    if (!e->profiling) return nullptr;
    if (d.prof_used + 5 > d.prof_events.size()) {
        size_t old = d.prof_events.size();
        d.prof_events.resize(old + 128);
        for (size_t i = old; i < d.prof_events.size(); i++)
            if (cudaEventCreate(&d.prof_events[i]) != cudaSuccess) { d.prof_events.resize(i); return nullptr; }
    }
    cudaEvent_t *ev = &d.prof_events[d.prof_used];
    d.prof_used += 5;  // start, after prep, before / after the fixed-base (or generic) kernel, after k_gpart
    return ev;
}

}  // namespace

void sbv_scratch_free(Dev &d) {
    for (auto &w : d.ws) {
        void *ptrs[] = {w.uw, w.flags, w.tscr, w.gacc, w.htab, w.rep, w.keylist, w.klist, w.glist, w.zeroed, w.keyid, w.item_kid, w.bases, w.hs, w.ztop, w.pref, w.ktab, w.keyflags};
        for (void *p : ptrs) if (p) cudaFree(p);
        cudaEvent_t evs[] = {w.done, w.ev_group, w.ev_prep, w.ev_tab, w.ev_gen};
        for (cudaEvent_t ev : evs) if (ev) cudaEventDestroy(ev);
        if (w.s_tab) cudaStreamDestroy(w.s_tab);
        if (w.s_gen) cudaStreamDestroy(w.s_gen);
        w = Dev::Scratch{};
    }
}

int sbv_init_gtables(sbv_engine *e, Dev &d) {
    for (int c = 0; c < 2; c++) {
        const CurveOps &ops = sbv_ops(c);
        CU(e, cudaMalloc(&d.gtab[c], ops.gtab_entries * 2 * ops.N * 4));  // P-256: 64 MiB, stays resident in the 126 MB L2
        CU(e, ops.gtable_init(d.gtab[c], d.stream));
    }
    e->launches += 2;
    CU(e, cudaStreamSynchronize(d.stream));
    return 0;
}

// The keys-per-item pipeline in two halves, so that a host-buffer call can upload the keys first and let the grouping
// and the table construction (the latency-bound part) run while the rest of the batch is still on its way:
//   begin : scratch set, key grouping on st, table construction on the set's side stream   (needs qx, qy)
//   finish: k_prep, generic kernel on the second side stream, fixed-base kernel, join       (needs r, s, digest)
int sbv_launch_verify_begin(sbv_engine *e, Dev &d, uint8_t curve, size_t n, const uint8_t *d_qx, const uint8_t *d_qy, cudaStream_t st,
                            VerifyLaunch *vl, int chunks) {
    *vl = VerifyLaunch{};
    if (n == 0) return 0;
    if (chunks < 1 || chunks > SBV_MAX_CHUNKS) return sbv_fail(e, SBV_ERR_ARG, "bad chunk count %d", chunks);
    const CurveOps &ops = sbv_ops(curve);
    const KtOps *kt = ops.kt5;
    const uint32_t nn = (uint32_t)n;
    const uint32_t T = e->group_threshold > 0 ? (uint32_t)e->group_threshold : 0;
    const bool grouping = T > 0 && n >= T && n >= (size_t)e->group_min_batch && e->group_max_keys > 0;
    size_t kcap = 0;
    if (grouping) {
        kcap = n / T;
        if (kcap > (size_t)e->group_max_keys) kcap = (size_t)e->group_max_keys;
        if (kcap == 0) kcap = 1;
    }
    Dev::Scratch *w = nullptr;
    if (int rc = take_scratch(e, d, ops, grouping ? kt : nullptr, n, kcap, st, &w)) return rc;
    w->open = true;  // until sbv_launch_verify_finish records the set's `done` event
    vl->w = w; vl->curve = curve; vl->n = n; vl->grouping = grouping; vl->d_qx = d_qx; vl->d_qy = d_qy; vl->chunks = chunks;
    vl->ev = prof_take(e, d);
    if (vl->ev) CU(e, cudaEventRecord(vl->ev[0], st));
    if (!grouping) return 0;
    uint32_t *counters = w->zeroed, *kcnt = w->zeroed + 4;
    CU(e, cudaMemsetAsync(w->htab, 0xff, (size_t)w->hsize * 4, st));
    CU(e, cudaMemsetAsync(w->zeroed, 0, (n + 4 + (chunks > 1 ? 4 * chunks : 0)) * 4, st));
    CU(e, ops.group(nn, d_qx, d_qy, e->hash_seed, w->hsize - 1, w->htab, w->rep, kcnt, T, (uint32_t)kcap, w->keyid, w->keylist, w->item_kid, w->klist,
                    w->glist, counters, chunks == 1, st));
    CU(e, cudaEventRecord(w->ev_group, st));
    CU(e, cudaStreamWaitEvent(w->s_tab, w->ev_group, 0));
    CU(e, kt->build(counters + 0, (uint32_t)kcap, w->keylist, d_qx, d_qy, w->bases, w->hs, w->ztop, w->pref, w->ktab, w->keyflags, w->s_tab));
    CU(e, cudaEventRecord(w->ev_tab, w->s_tab));
    e->launches += chunks == 1 ? 7 : 6;
    return 0;
}

// One chunk of a chunked launch: the items [lo, lo + cn) are a batch of their own as far as the per-item arrays go (every one
// of them is word-major with the batch size as its stride, so the chunk's slice is the contiguous block at `words per item *
// lo`); what the chunks share is the grouping (hash table, rep, key ids) and the key tables.
int sbv_launch_verify_chunk(sbv_engine *e, Dev &d, const VerifyLaunch &vl, int c, size_t lo, size_t cn, bool last, const uint8_t *d_r, const uint8_t *d_s,
                            const uint8_t *d_dig, uint32_t dlen, uint8_t *d_ok, cudaStream_t st) {
    if (vl.n == 0) return 0;
    Dev::Scratch *w = vl.w;
    if (vl.chunks <= 1 || c < 0 || c >= vl.chunks || lo + cn > vl.n) return sbv_fail(e, SBV_ERR_ARG, "bad chunk");
    const CurveOps &ops = sbv_ops(vl.curve);
    const KtOps *kt = ops.kt5;
    const size_t N = (size_t)ops.N, L = (size_t)ops.bytes;
    const uint32_t nn = (uint32_t)cn;
    const uint32_t *gtab = d.gtab[vl.curve];
    cudaEvent_t *ev = last ? vl.ev : nullptr;  // the profile of a chunked launch is that of its last chunk (nothing of the first half overlaps it)
    if (ev) CU(e, cudaEventRecord(ev[0], st));
    uint32_t *uw = w->uw + 2 * N * lo, *tscr = w->tscr + 12 * N * lo;
    uint8_t *flags = w->flags + lo;
    const uint8_t *r = d_r + lo * L;
    if (cn) {
        CU(e, ops.prep(nn, r, d_s + lo * L, d_dig + lo * dlen, dlen, uw, flags, st));
        e->launches += 1;
    }
    if (ev) CU(e, cudaEventRecord(ev[1], st));
    if (!vl.grouping) {
        if (ev) { CU(e, cudaEventRecord(ev[4], st)); CU(e, cudaEventRecord(ev[2], st)); }
        if (cn) {
            CU(e, ops.coz(nn, vl.d_qx + lo * L, vl.d_qy + lo * L, r, uw, flags, gtab, tscr, d_ok + lo, nullptr, nullptr, st));
            e->launches += 1;
        }
        if (ev) CU(e, cudaEventRecord(ev[3], st));
    } else if (cn) {
        uint32_t *cc = w->zeroed + 4 + vl.n + 4 * (size_t)c;   // this chunk's counters (zeroed by the first half)
        uint32_t *klist = w->klist + lo, *glist = w->glist + lo;
        CU(e, ops.route(nn, w->rep + lo, w->keyid, w->item_kid + lo, klist, glist, cc, st));
        CU(e, cudaEventRecord(w->ev_prep, st));
        CU(e, cudaStreamWaitEvent(w->s_gen, w->ev_prep, 0));
        CU(e, ops.coz(nn, vl.d_qx + lo * L, vl.d_qy + lo * L, r, uw, flags, gtab, tscr, d_ok + lo, glist, cc + 2, w->s_gen));
        CU(e, cudaEventRecord(w->ev_gen, w->s_gen));
        const uint32_t *gacc = nullptr;
        if (e->gsplit) {
            uint32_t *ga = w->gacc + 3 * N * lo;
            CU(e, ops.gpart(nn, uw, gtab, ga, st));
            gacc = ga;
            e->launches += 1;
        }
        if (ev) CU(e, cudaEventRecord(ev[4], st));
        if (c == 0) CU(e, cudaStreamWaitEvent(st, w->ev_tab, 0));
        if (ev) CU(e, cudaEventRecord(ev[2], st));
        CU(e, kt->verify(0, 0, nn, nullptr, w->item_kid + lo, 0, w->keyflags, r, uw, flags, gtab, w->ktab, d_ok + lo, klist, cc + 1, gacc, st));
        if (ev) CU(e, cudaEventRecord(ev[3], st));
        CU(e, cudaStreamWaitEvent(st, w->ev_gen, 0));
        e->launches += 3;
    } else if (ev) {
        CU(e, cudaEventRecord(ev[4], st)); CU(e, cudaEventRecord(ev[2], st)); CU(e, cudaEventRecord(ev[3], st));
    }
    if (last) {
        CU(e, cudaEventRecord(w->done, st));
        w->open = false;
    }
    return 0;
}

// A fault between the halves: the set goes back, but only behind whatever the first half left running on its side stream.
void sbv_launch_verify_abort(const VerifyLaunch &vl, cudaStream_t st) {
    Dev::Scratch *w = vl.w;
    if (!w || !w->open) return;
    if (vl.grouping) cudaStreamWaitEvent(st, w->ev_tab, 0);
    cudaStreamWaitEvent(st, w->ev_gen, 0);   // a chunk's generic kernel (a never-recorded event is a no-op)
    cudaEventRecord(w->done, st);
    w->open = false;
}

int sbv_launch_verify_finish(sbv_engine *e, Dev &d, const VerifyLaunch &vl, const uint8_t *d_r, const uint8_t *d_s, const uint8_t *d_dig,
                             uint32_t dlen, uint8_t *d_ok, cudaStream_t st) {
    if (vl.n == 0) return 0;
    if (vl.chunks != 1) return sbv_fail(e, SBV_ERR_ARG, "chunked launch finished in one piece");
    const CurveOps &ops = sbv_ops(vl.curve);
    const KtOps *kt = ops.kt5;
    Dev::Scratch *w = vl.w;
    cudaEvent_t *ev = vl.ev;
    const uint32_t nn = (uint32_t)vl.n;
    const uint32_t *gtab = d.gtab[vl.curve];
    CU(e, ops.prep(nn, d_r, d_s, d_dig, dlen, w->uw, w->flags, st));
    if (!vl.grouping) {
        if (ev) { CU(e, cudaEventRecord(ev[1], st)); CU(e, cudaEventRecord(ev[4], st)); CU(e, cudaEventRecord(ev[2], st)); }
        CU(e, ops.coz(nn, vl.d_qx, vl.d_qy, d_r, w->uw, w->flags, gtab, w->tscr, d_ok, nullptr, nullptr, st));
        if (ev) CU(e, cudaEventRecord(ev[3], st));
        CU(e, cudaEventRecord(w->done, st));
        w->open = false;
        e->launches += 2;
        return 0;
    }
    uint32_t *counters = w->zeroed;
    CU(e, cudaEventRecord(w->ev_prep, st));
    if (ev) CU(e, cudaEventRecord(ev[1], st));
    CU(e, cudaStreamWaitEvent(w->s_gen, w->ev_prep, 0));
    CU(e, ops.coz(nn, vl.d_qx, vl.d_qy, d_r, w->uw, w->flags, gtab, w->tscr, d_ok, w->glist, counters + 2, w->s_gen));
    CU(e, cudaEventRecord(w->ev_gen, w->s_gen));
    // the u1*G half needs no table: it runs while the tables are still being built
    const uint32_t *gacc = nullptr;
    if (e->gsplit) {
        CU(e, ops.gpart(nn, w->uw, gtab, w->gacc, st));
        gacc = w->gacc;
        e->launches += 1;
    }
    if (ev) CU(e, cudaEventRecord(ev[4], st));  // == ev[1] without the split
    CU(e, cudaStreamWaitEvent(st, w->ev_tab, 0));
    if (ev) CU(e, cudaEventRecord(ev[2], st));
    CU(e, kt->verify(0, 0, nn, nullptr, w->item_kid, 0, w->keyflags, d_r, w->uw, w->flags, gtab, w->ktab, d_ok, w->klist, counters + 1, gacc, st));
    if (ev) CU(e, cudaEventRecord(ev[3], st));
    CU(e, cudaStreamWaitEvent(st, w->ev_gen, 0));
    CU(e, cudaEventRecord(w->done, st));
    w->open = false;
    e->launches += 3;
    return 0;
}

int sbv_launch_verify(sbv_engine *e, Dev &d, uint8_t curve, size_t n, const uint8_t *d_r, const uint8_t *d_s, const uint8_t *d_qx,
                      const uint8_t *d_qy, const uint8_t *d_dig, uint32_t dlen, uint8_t *d_ok, cudaStream_t st) {
    VerifyLaunch vl;
    if (int rc = sbv_launch_verify_begin(e, d, curve, n, d_qx, d_qy, st, &vl)) return rc;
    return sbv_launch_verify_finish(e, d, vl, d_r, d_s, d_dig, dlen, d_ok, st);
}

// ---- registered keys ----------------------------------------------------------------------------------------
void sbv_keys_free(Dev &d) {
    for (int c = 0; c < 2; c++) {
        if (d.ktab[c]) cudaFree(d.ktab[c]);
        if (d.keyflags[c]) cudaFree(d.keyflags[c]);
        if (d.slot2local[c]) cudaFree(d.slot2local[c]);
        d.ktab[c] = nullptr; d.keyflags[c] = nullptr; d.slot2local[c] = nullptr; d.n_local[c] = 0;
    }
    d.n_slots = 0;
}

// Consenter keys are configuration (they change only with a reconfiguration, i.e. a new VerificationSequence —
// /root/reference/pkg/api/dependencies.go:65-66): one table per key with 8-bit signed windows, built by the same
// four kernels the on-the-fly path uses (a few milliseconds for a thousand keys).
int sbv_keys_build(sbv_engine *e, Dev &d) {
    CU(e, cudaSetDevice(d.ordinal));
    CU(e, cudaDeviceSynchronize());  // no launch on any lane may still read the old tables
    sbv_keys_free(d);
    const size_t n = e->key_ids.size();
    d.n_slots = (uint32_t)n;
    if (n == 0) return 0;
    for (int c = 0; c < 2; c++) {
        const CurveOps &ops = sbv_ops(c);
        const KtOps *kt = ops.kt8;
        const size_t L = (size_t)ops.bytes;
        std::vector<int32_t> map(n, -1);
        std::vector<uint8_t> kx, ky;
        uint32_t cnt = 0;
        for (size_t i = 0; i < n; i++) {
            if (e->key_curve[i] != c) continue;
            const uint8_t *x = &e->key_xy[96 * i], *y = x + 48;
            bool fits = true;
            for (size_t b = 0; b < 48 - L; b++) if (x[b] || y[b]) fits = false;
            if (!fits) continue;  // value >= 2^(8L): not a valid key for this curve -> slot stays unmapped (rejects)
            map[i] = (int32_t)cnt++;
            kx.insert(kx.end(), x + (48 - L), x + 48);
            ky.insert(ky.end(), y + (48 - L), y + 48);
        }
        CU(e, cudaMalloc(&d.slot2local[c], n * sizeof(int32_t)));
        CU(e, cudaMemcpyAsync(d.slot2local[c], map.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, d.stream));
        d.n_local[c] = cnt;
        if (cnt == 0) { CU(e, cudaStreamSynchronize(d.stream)); continue; }
        CU(e, cudaMalloc(&d.ktab[c], kt->geom.ktab_words * cnt * 4));
        CU(e, cudaMalloc(&d.keyflags[c], cnt));
        uint8_t *d_kx = nullptr, *d_ky = nullptr;
        uint32_t *tmp = nullptr, *d_cnt = nullptr;
        const size_t tmp_words = (kt->geom.bases_words + kt->geom.hs_words + 2 * kt->geom.ztop_words) * cnt;
        CU(e, cudaMalloc(&d_kx, kx.size()));
        CU(e, cudaMalloc(&d_ky, ky.size()));
        CU(e, cudaMalloc(&tmp, tmp_words * 4));
        CU(e, cudaMalloc(&d_cnt, 4));
        CU(e, cudaMemcpyAsync(d_kx, kx.data(), kx.size(), cudaMemcpyHostToDevice, d.stream));
        CU(e, cudaMemcpyAsync(d_ky, ky.data(), ky.size(), cudaMemcpyHostToDevice, d.stream));
        CU(e, cudaMemcpyAsync(d_cnt, &cnt, 4, cudaMemcpyHostToDevice, d.stream));
        uint32_t *bases = tmp, *hs = bases + kt->geom.bases_words * cnt, *ztop = hs + kt->geom.hs_words * cnt, *pref = ztop + kt->geom.ztop_words * cnt;
        cudaError_t st = kt->build(d_cnt, cnt, nullptr, d_kx, d_ky, bases, hs, ztop, pref, d.ktab[c], d.keyflags[c], d.stream);
        e->launches += 4;
        if (st == cudaSuccess) st = cudaStreamSynchronize(d.stream);
        cudaFree(d_kx); cudaFree(d_ky); cudaFree(tmp); cudaFree(d_cnt);
        CU(e, st);
    }
    return 0;
}

int sbv_launch_keyed(sbv_engine *e, Dev &d, uint8_t curve, size_t n, const uint32_t *d_slot, const uint8_t *d_r, const uint8_t *d_s,
                     const uint8_t *d_dig, uint32_t dlen, uint8_t *d_ok, cudaStream_t st) {
    if (n == 0) return 0;
    if (d.n_local[curve] == 0) {  // no registered key of this curve: every item rejects
        CU(e, cudaMemsetAsync(d_ok, 0, n, st));
        return 0;
    }
    const CurveOps &ops = sbv_ops(curve);
    const KtOps *kt = ops.kt8;
    const uint32_t nn = (uint32_t)n;
    Dev::Scratch *w = nullptr;
    if (int rc = take_scratch(e, d, ops, nullptr, n, 0, st, &w)) return rc;
    cudaEvent_t *ev = prof_take(e, d);
    if (ev) CU(e, cudaEventRecord(ev[0], st));
    CU(e, ops.prep(nn, d_r, d_s, d_dig, dlen, w->uw, w->flags, st));
    if (ev) { CU(e, cudaEventRecord(ev[1], st)); CU(e, cudaEventRecord(ev[4], st)); CU(e, cudaEventRecord(ev[2], st)); }
    const int warp = nn <= (uint32_t)e->keyed_warp_limit ? 1 : 0;  // small batch: one signature per warp (latency path)
    CU(e, kt->verify(1, warp, nn, d_slot, d.slot2local[curve], d.n_slots, d.keyflags[curve], d_r, w->uw, w->flags, d.gtab[curve], d.ktab[curve], d_ok,
                     nullptr, nullptr, nullptr, st));
    if (ev) CU(e, cudaEventRecord(ev[3], st));
    CU(e, cudaEventRecord(w->done, st));
    e->launches += 2;
    return 0;
}
