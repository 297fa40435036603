// keyed.cu — registered-key path: per-key comb tables built at sbv_set_keys, fixed-base verification.
//
// Consenter keys are configuration (they change only with a reconfiguration, i.e. a new
// VerificationSequence — /root/reference/pkg/api/dependencies.go:65-66), so the engine precomputes
// K[k][i][b] = b * 2^(8i) * Q_k once per key; verification is then 2*BYTES mixed additions.
#include "engine.h"
#include "kernels.cuh"

using namespace sbv;

void sbv_keys_free(Dev &d) {
    for (int c = 0; c < 2; c++) {
        if (d.ktab[c]) cudaFree(d.ktab[c]);
        if (d.keyflags[c]) cudaFree(d.keyflags[c]);
        if (d.slot2local[c]) cudaFree(d.slot2local[c]);
        d.ktab[c] = nullptr; d.keyflags[c] = nullptr; d.slot2local[c] = nullptr; d.n_local[c] = 0;
    }
    d.n_slots = 0;
}

int sbv_keys_build(sbv_engine *e, Dev &d) {
    CU(e, cudaSetDevice(d.ordinal));
    CU(e, cudaDeviceSynchronize());  // no launch on any lane may still read the old tables
    sbv_keys_free(d);
    const size_t n = e->key_ids.size();
    d.n_slots = (uint32_t)n;
    if (n == 0) return 0;
    for (int c = 0; c < 2; c++) {
        const size_t L = c == 0 ? 32 : 48, N = c == 0 ? 8 : 12;
        std::vector<int32_t> map(n, -1);
        std::vector<uint8_t> keys;
        uint32_t cnt = 0;
        for (size_t i = 0; i < n; i++) {
            if (e->key_curve[i] != c) continue;
            const uint8_t *x = &e->key_xy[96 * i], *y = x + 48;
            bool fits = true;
            for (size_t b = 0; b < 48 - L; b++) if (x[b] || y[b]) fits = false;
            if (!fits) continue;  // value >= 2^(8L): not a valid key for this curve -> slot stays unmapped (rejects)
            map[i] = (int32_t)cnt++;
            keys.insert(keys.end(), x + (48 - L), x + 48);
            keys.insert(keys.end(), y + (48 - L), y + 48);
        }
        CU(e, cudaMalloc(&d.slot2local[c], n * sizeof(int32_t)));
        CU(e, cudaMemcpyAsync(d.slot2local[c], map.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, d.stream));
        d.n_local[c] = cnt;
        if (cnt == 0) { CU(e, cudaStreamSynchronize(d.stream)); continue; }
        const size_t entries = (size_t)cnt * L * 256;
        CU(e, cudaMalloc(&d.ktab[c], entries * 2 * N * 4));
        CU(e, cudaMalloc(&d.keyflags[c], cnt));
        uint8_t *d_keys = nullptr;
        CU(e, cudaMalloc(&d_keys, keys.size()));
        CU(e, cudaMemcpyAsync(d_keys, keys.data(), keys.size(), cudaMemcpyHostToDevice, d.stream));
        const uint32_t blocks = (uint32_t)((entries + 127) / 128);
        if (c == 0) k_keytab_init<P256><<<blocks, 128, 0, d.stream>>>(cnt, d_keys, d.ktab[c], d.keyflags[c]);
        else k_keytab_init<P384><<<blocks, 128, 0, d.stream>>>(cnt, d_keys, d.ktab[c], d.keyflags[c]);
        e->launches += 1;
        CU(e, cudaGetLastError());
        CU(e, cudaStreamSynchronize(d.stream));
        cudaFree(d_keys);
    }
    return 0;
}

template <class C, int BLOCK, int MINB>
static int launch_keyed_t(sbv_engine *e, Dev &d, int c, size_t n, const uint32_t *d_slot, const uint8_t *d_r, const uint8_t *d_s,
                          const uint8_t *d_dig, uint32_t dlen, uint8_t *d_ok, cudaStream_t st) {
    constexpr int S = 8;
    const uint32_t nn = (uint32_t)n;
    Dev::Scratch *w = nullptr;
    if (int rc = sbv_take_scratch(e, d, st, &w)) return rc;
    cudaEvent_t *ev = nullptr;
    if (e->profiling) {
        if (d.prof_used + 3 > d.prof_events.size()) {
            size_t old = d.prof_events.size();
            d.prof_events.resize(old + 96);
            for (size_t i = old; i < d.prof_events.size(); i++) CU(e, cudaEventCreate(&d.prof_events[i]));
        }
        ev = &d.prof_events[d.prof_used];
        d.prof_used += 3;
        CU(e, cudaEventRecord(ev[0], st));
    }
    CU(e, (launch_prep<C, 0, S>(nn, d_r, d_s, d_dig, dlen, w->gidx, w->digits, w->flags, st)));
    if (ev) CU(e, cudaEventRecord(ev[1], st));
    const uint32_t warp_limit = (uint32_t)e->keyed_warp_limit;
    if (nn <= warp_limit)  // small batch: one signature per warp (latency path)
        k_verify_keyed_warp<C><<<(nn * 32 + 127) / 128, 128, 0, st>>>(
            nn, d_slot, d.slot2local[c], d.n_slots, d.keyflags[c], d_r, w->gidx, reinterpret_cast<const uint8_t *>(w->digits), w->flags,
            reinterpret_cast<const uint4 *>(d.gtab[c]), reinterpret_cast<const uint4 *>(d.ktab[c]), d_ok);
    else
    k_verify_keyed<C, BLOCK, MINB><<<(nn + BLOCK - 1) / BLOCK, BLOCK, 0, st>>>(
        nn, d_slot, d.slot2local[c], d.n_slots, d.keyflags[c], d_r, w->gidx, reinterpret_cast<const uint8_t *>(w->digits), w->flags,
        reinterpret_cast<const uint4 *>(d.gtab[c]), reinterpret_cast<const uint4 *>(d.ktab[c]), d_ok);
    if (ev) CU(e, cudaEventRecord(ev[2], st));
    CU(e, cudaEventRecord(w->done, st));
    e->launches += 2;
    CU(e, cudaGetLastError());
    return 0;
}

int sbv_launch_keyed(sbv_engine *e, Dev &d, uint8_t curve, size_t n, const uint32_t *d_slot, const uint8_t *d_r, const uint8_t *d_s,
                     const uint8_t *d_dig, uint32_t dlen, uint8_t *d_ok, cudaStream_t st) {
    if (n == 0) return 0;
    if (d.n_local[curve] == 0) {  // no registered key of this curve: every item rejects
        CU(e, cudaMemsetAsync(d_ok, 0, n, st));
        return 0;
    }
    // P-256: 7 blocks of 64 threads per SM = 66,304 resident threads >= one 65,536 batch (single wave)
    if (curve == SBV_P256) return launch_keyed_t<P256, 64, 7>(e, d, 0, n, d_slot, d_r, d_s, d_dig, dlen, d_ok, st);
    return launch_keyed_t<P384, 64, 4>(e, d, 1, n, d_slot, d_r, d_s, d_dig, dlen, d_ok, st);
}

extern "C" {

int sbv_verify_registered_device(sbv_engine *e, int device_index, uint8_t curve, size_t n, const uint32_t *d_key_slot,
                                 const uint8_t *d_r, const uint8_t *d_s, const uint8_t *d_digest, uint8_t digest_len, uint8_t *d_ok,
                                 void *cuda_stream) {
    if (!e || curve > SBV_P384 || device_index < 0 || device_index >= (int)e->devs.size() || digest_len == 0 || (digest_len & 3))
        return sbv_fail(e, SBV_ERR_ARG, "sbv_verify_registered_device: bad argument");
    if (n == 0) return SBV_OK;
    if (n > 0x7fffffffu) return sbv_fail(e, SBV_ERR_ARG, "n too large");
    std::lock_guard<std::mutex> lk(e->mu);
    Dev &d = e->devs[device_index];
    CU(e, cudaSetDevice(d.ordinal));
    int rc = sbv_ensure_workspace(e, d, n);
    if (rc) return rc;
    return sbv_launch_keyed(e, d, curve, n, d_key_slot, d_r, d_s, d_digest, digest_len, d_ok, (cudaStream_t)cuda_stream);
}

int sbv_verify_registered(sbv_engine *e, uint8_t curve, size_t n, const uint32_t *key_slot, const uint8_t *r, const uint8_t *s,
                          const uint8_t *digest, uint8_t digest_len, uint8_t *ok) {
    if (!e || curve > SBV_P384 || digest_len == 0 || (digest_len & 3) || digest_len > 64)
        return sbv_fail(e, SBV_ERR_ARG, "sbv_verify_registered: bad argument");
    if (n == 0) return SBV_OK;
    if (!key_slot || !r || !s || !digest || !ok) return sbv_fail(e, SBV_ERR_ARG, "null buffer");
    if (n > 0x7fffffffu) return sbv_fail(e, SBV_ERR_ARG, "n too large");
    const size_t L = curve == SBV_P256 ? 32 : 48;
    const int G = (int)e->devs.size();
    const int lane = sbv_lane_acquire(e);
    struct Release { sbv_engine *e; int lane; ~Release() { sbv_lane_release(e, lane); } } release{e, lane};
    for (int g = 0; g < G; g++) {
        Dev &d = e->devs[g];
        Dev::Lane &ln = d.lanes[lane];
        const size_t lo = n * g / G, cnt = n * (g + 1) / G - lo;
        if (cnt == 0) continue;
        CU(e, cudaSetDevice(d.ordinal));
        int rc = sbv_lane_ensure(e, d, ln, cnt, cnt * (2 * L + digest_len + 5) + 8 * 256);
        if (rc) return rc;
        size_t so = 0;
        if ((rc = sbv_lane_h2d(e, ln, ln.d_slot, key_slot + lo, cnt * 4, so))) return rc;
        if ((rc = sbv_lane_h2d(e, ln, ln.d_r, r + lo * L, cnt * L, so))) return rc;
        if ((rc = sbv_lane_h2d(e, ln, ln.d_s, s + lo * L, cnt * L, so))) return rc;
        if ((rc = sbv_lane_h2d(e, ln, ln.d_dig, digest + lo * digest_len, cnt * digest_len, so))) return rc;
        {
            std::lock_guard<std::mutex> lk(e->mu);
            rc = sbv_ensure_workspace(e, d, cnt);
            if (!rc) rc = sbv_launch_keyed(e, d, curve, cnt, ln.d_slot, ln.d_r, ln.d_s, ln.d_dig, digest_len, ln.d_ok, ln.stream);
        }
        if (rc) return rc;
        CU(e, cudaMemcpyAsync(ok + lo, ln.d_ok, cnt, cudaMemcpyDeviceToHost, ln.stream));
    }
    for (int g = 0; g < G; g++) {
        CU(e, cudaSetDevice(e->devs[g].ordinal));
        if (e->devs[g].lanes[lane].stream) CU(e, cudaStreamSynchronize(e->devs[g].lanes[lane].stream));
    }
    return SBV_OK;
}

// Fused SHA-256 -> registered-key verify: messages hashed on the device, digests never leave it.
int sbv_hash_verify_registered(sbv_engine *e, uint8_t curve, size_t n, const uint8_t *msgs, const uint64_t *msg_off,
                               const uint32_t *key_slot, const uint8_t *r, const uint8_t *s, uint8_t *ok) {
    if (!e || curve > SBV_P384) return sbv_fail(e, SBV_ERR_ARG, "sbv_hash_verify_registered: bad argument");
    if (n == 0) return SBV_OK;
    if (!msg_off || !key_slot || !r || !s || !ok || (!msgs && msg_off[n] != msg_off[0])) return sbv_fail(e, SBV_ERR_ARG, "null buffer");
    if (n > 0x7fffffffu) return sbv_fail(e, SBV_ERR_ARG, "n too large");
    const size_t L = curve == SBV_P256 ? 32 : 48;
    const int G = (int)e->devs.size();
    const int lane = sbv_lane_acquire(e);
    struct Release { sbv_engine *e; int lane; ~Release() { sbv_lane_release(e, lane); } } release{e, lane};
    for (int g = 0; g < G; g++) {
        Dev &d = e->devs[g];
        Dev::Lane &ln = d.lanes[lane];
        const size_t lo = n * g / G, cnt = n * (g + 1) / G - lo;
        if (cnt == 0) continue;
        CU(e, cudaSetDevice(d.ordinal));
        const uint64_t bytes = msg_off[lo + cnt] - msg_off[lo];
        int rc = sbv_lane_ensure(e, d, ln, cnt, cnt * (2 * L + 5 + 8) + bytes + 16 * 256);
        if (rc) return rc;
        rc = sbv_lane_ensure_msgs(e, ln, bytes + 16, cnt + 1);
        if (rc) return rc;
        size_t so = 0;
        if ((rc = sbv_lane_h2d(e, ln, ln.d_msgs, msgs + msg_off[lo], bytes, so))) return rc;
        if ((rc = sbv_lane_h2d(e, ln, ln.d_off, msg_off + lo, (cnt + 1) * 8, so))) return rc;
        if ((rc = sbv_lane_h2d(e, ln, ln.d_slot, key_slot + lo, cnt * 4, so))) return rc;
        if ((rc = sbv_lane_h2d(e, ln, ln.d_r, r + lo * L, cnt * L, so))) return rc;
        if ((rc = sbv_lane_h2d(e, ln, ln.d_s, s + lo * L, cnt * L, so))) return rc;
        {
            std::lock_guard<std::mutex> lk(e->mu);
            rc = sbv_launch_sha256(e, cnt, ln.d_msgs, ln.d_off, msg_off[lo], ln.d_dig, ln.d_perm, ln.stream);
            if (!rc) rc = sbv_ensure_workspace(e, d, cnt);
            if (!rc) rc = sbv_launch_keyed(e, d, curve, cnt, ln.d_slot, ln.d_r, ln.d_s, ln.d_dig, 32, ln.d_ok, ln.stream);
        }
        if (rc) return rc;
        CU(e, cudaMemcpyAsync(ok + lo, ln.d_ok, cnt, cudaMemcpyDeviceToHost, ln.stream));
    }
    for (int g = 0; g < G; g++) {
        CU(e, cudaSetDevice(e->devs[g].ordinal));
        if (e->devs[g].lanes[lane].stream) CU(e, cudaStreamSynchronize(e->devs[g].lanes[lane].stream));
    }
    return SBV_OK;
}

}  // extern "C"
