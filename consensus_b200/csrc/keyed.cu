// keyed.cu — registered-key entry points (sbv_verify_registered*, sbv_hash_verify_registered).
//
// Consenter keys are configuration (they change only with a reconfiguration, i.e. a new VerificationSequence —
// /root/reference/pkg/api/dependencies.go:65-66): sbv_set_keys builds one fixed-base table per key (pipeline.cu:
// sbv_keys_build), and verification against a registered key is NWIN + GWINS mixed additions with no doublings.
#include "engine.h"

extern "C" {

int sbv_verify_registered_device(sbv_engine *e, int device_index, uint8_t curve, size_t n, const uint32_t *d_key_slot,
                                 const uint8_t *d_r, const uint8_t *d_s, const uint8_t *d_digest, uint8_t digest_len, uint8_t *d_ok,
                                 void *cuda_stream) {
    if (!e || curve > SBV_P384 || device_index < 0 || device_index >= (int)e->devs.size() || digest_len == 0 || (digest_len & 3) || digest_len > 64)
        return sbv_fail(e, SBV_ERR_ARG, "sbv_verify_registered_device: bad argument");
    if (n == 0) return SBV_OK;
    if (n > 0x7fffffffu) return sbv_fail(e, SBV_ERR_ARG, "n too large");
    if (!d_key_slot || !d_r || !d_s || !d_digest || !d_ok) return sbv_fail(e, SBV_ERR_ARG, "null buffer");
    std::lock_guard<std::mutex> lk(e->mu);
    Dev &d = e->devs[device_index];
    CU(e, cudaSetDevice(d.ordinal));
    return sbv_launch_keyed(e, d, curve, n, d_key_slot, d_r, d_s, d_digest, digest_len, d_ok, (cudaStream_t)cuda_stream);
}

static int sync_lane(sbv_engine *e, int lane) {
    for (Dev &d : e->devs) {
        if (!d.lanes[lane].stream) continue;
        CU(e, cudaSetDevice(d.ordinal));
        CU(e, cudaStreamSynchronize(d.lanes[lane].stream));
    }
    return 0;
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
    LaneGuard guard(e);
    const int lane = guard.lane;
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
            rc = sbv_launch_keyed(e, d, curve, cnt, ln.d_slot, ln.d_r, ln.d_s, ln.d_dig, digest_len, ln.d_ok, ln.stream);
        }
        if (rc) return rc;
        CU(e, cudaMemcpyAsync(ok + lo, ln.d_ok, cnt, cudaMemcpyDeviceToHost, ln.stream));
    }
    return sync_lane(e, lane);
}

// Fused SHA-256 -> registered-key verify: messages hashed on the device, digests never leave it.
int sbv_hash_verify_registered(sbv_engine *e, uint8_t curve, size_t n, const uint8_t *msgs, const uint64_t *msg_off,
                               const uint32_t *key_slot, const uint8_t *r, const uint8_t *s, uint8_t *ok) {
    if (!e || curve > SBV_P384) return sbv_fail(e, SBV_ERR_ARG, "sbv_hash_verify_registered: bad argument");
    if (n == 0) return SBV_OK;
    if (!msg_off || !key_slot || !r || !s || !ok || (!msgs && msg_off[n] != msg_off[0])) return sbv_fail(e, SBV_ERR_ARG, "null buffer");
    if (n > 0x7fffffffu) return sbv_fail(e, SBV_ERR_ARG, "n too large");
    for (size_t i = 0; i < n; i++) if (msg_off[i + 1] < msg_off[i]) return sbv_fail(e, SBV_ERR_ARG, "msg_off is not non-decreasing");
    const size_t L = curve == SBV_P256 ? 32 : 48;
    const int G = (int)e->devs.size();
    LaneGuard guard(e);
    const int lane = guard.lane;
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
        if ((rc = sbv_launch_sha256(e, cnt, ln.d_msgs, ln.d_off, msg_off[lo], ln.d_dig, ln.d_perm, ln.stream))) return rc;
        {
            std::lock_guard<std::mutex> lk(e->mu);
            rc = sbv_launch_keyed(e, d, curve, cnt, ln.d_slot, ln.d_r, ln.d_s, ln.d_dig, 32, ln.d_ok, ln.stream);
        }
        if (rc) return rc;
        CU(e, cudaMemcpyAsync(ok + lo, ln.d_ok, cnt, cudaMemcpyDeviceToHost, ln.stream));
    }
    return sync_lane(e, lane);
}

}  // extern "C"
