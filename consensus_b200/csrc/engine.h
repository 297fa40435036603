// engine.h — shared host-side state of libsbv.so (one engine = 1..8 devices of one box).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdint>
#include <condition_variable>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/sbv.h"

struct Dev {
    int ordinal = 0;
    cudaStream_t stream = nullptr;
    uint32_t *gtab[2] = {nullptr, nullptr};
    // per-batch workspace (device)
    size_t cap = 0;
    uint8_t *d_r = nullptr, *d_s = nullptr, *d_qx = nullptr, *d_qy = nullptr, *d_dig = nullptr, *d_ok = nullptr;
    // k_prep -> k_verify scratch, double-buffered so that launches on different streams may overlap:
    // a launch takes the next set and first waits for the event of that set's previous user.
    struct Scratch {
        uint16_t *gidx = nullptr;
        int8_t *digits = nullptr;
        uint8_t *flags = nullptr;
        uint32_t *tscr = nullptr;  // k_verify_coz per-signature scratch: 12N words, word-major
        cudaEvent_t done = nullptr;
        bool used = false;
    } ws[4];
    unsigned ws_next = 0;
    // per-call lanes of the host-buffer entry points: own stream, input/verdict buffers and pinned staging, so
    // two host threads can have a call in flight each (H2D / kernels / D2H of one overlap the other's)
    struct Lane {
        cudaStream_t stream = nullptr;
        size_t cap = 0;
        uint8_t *d_r = nullptr, *d_s = nullptr, *d_qx = nullptr, *d_qy = nullptr, *d_dig = nullptr, *d_ok = nullptr;
        uint32_t *d_slot = nullptr;
        uint8_t *d_msgs = nullptr;
        uint64_t *d_off = nullptr;
        uint32_t *d_perm = nullptr;  // off_cap entries + 3 * 1024 words of sort state
        size_t msg_cap = 0, off_cap = 0;
        uint8_t *h_pin = nullptr;
        size_t h_pin_cap = 0;
    } lanes[2];
    // message workspace
    size_t msg_cap = 0, off_cap = 0;
    uint8_t *d_msgs = nullptr;
    uint64_t *d_off = nullptr;
    uint32_t *d_perm = nullptr;
    // pinned staging
    uint8_t *h_pin = nullptr;
    size_t h_pin_cap = 0;
    // generic scratch (quorum, bitmask)
    uint8_t *d_scratch = nullptr;
    size_t scratch_cap = 0;
    // registered keys (sbv_set_keys): per-curve comb tables, validity flags, slot -> table index
    uint32_t *ktab[2] = {nullptr, nullptr};
    uint8_t *keyflags[2] = {nullptr, nullptr};
    int32_t *slot2local[2] = {nullptr, nullptr};
    uint32_t n_slots = 0, n_local[2] = {0, 0};
    // profiling: event pairs around the prep / verify kernels (only when enabled)
    std::vector<cudaEvent_t> prof_events;  // triples: before prep, between, after verify
    size_t prof_used = 0;
};

struct sbv_engine {
    std::vector<Dev> devs;
    std::mutex mu;
    std::condition_variable lane_cv;
    bool lane_busy[2] = {false, false};
    std::string err;
    uint64_t launches = 0;
    int keyed_warp_limit = 2048;  // registered-key batches up to this size use one warp per signature (SBV_KEYED_WARP_LIMIT)
    int p384_variant = 1;  // 1 = co-Z 4-bit window, 0 = 3-bit Jacobian window (SBV_P384_VARIANT)
    int p256_variant = 1;  // 1 = co-Z 4-bit window (default), 2 = co-Z with one lockstep 448-thread block per SM, 0 = 3-bit Jacobian window
    bool profiling = false;
    // NCCL (multi-device only; loaded lazily with dlopen so single-device users never touch it)
    void *nccl_lib = nullptr;
    std::vector<void *> nccl_comms;
    std::vector<uint32_t> gather_words;  // host copy of the gathered verdict bitmask
    // key registry
    uint64_t verification_seq = 0;
    std::vector<uint64_t> key_ids;
    std::vector<uint8_t> key_curve;
    std::vector<uint8_t> key_xy;
};

inline int sbv_fail(sbv_engine *e, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (e) e->err = buf;
    return code;
}
#define fail sbv_fail

#define CU(e, call)                                                                                   \
    do {                                                                                              \
        cudaError_t _st = (call);                                                                     \
        if (_st != cudaSuccess)                                                                       \
            return sbv_fail(e, SBV_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_st), \
                            __FILE__, __LINE__);                                                      \
    } while (0)

// per-(curve, window, block) kernel launchers — one translation unit each (inst_*.cu)
int sbv_launch_p256_w3_b64(sbv_engine *e, Dev &d, size_t n, const uint8_t *d_r, const uint8_t *d_s, const uint8_t *d_qx,
                           const uint8_t *d_qy, const uint8_t *d_dig, uint32_t dlen, uint8_t *d_ok, cudaStream_t st);
int sbv_launch_p256_coz_b64(sbv_engine *e, Dev &d, size_t n, const uint8_t *d_r, const uint8_t *d_s, const uint8_t *d_qx,
                            const uint8_t *d_qy, const uint8_t *d_dig, uint32_t dlen, uint8_t *d_ok, cudaStream_t st);
int sbv_launch_p256_coz_b448(sbv_engine *e, Dev &d, size_t n, const uint8_t *d_r, const uint8_t *d_s, const uint8_t *d_qx,
                             const uint8_t *d_qy, const uint8_t *d_dig, uint32_t dlen, uint8_t *d_ok, cudaStream_t st);
int sbv_launch_p384_coz_b64(sbv_engine *e, Dev &d, size_t n, const uint8_t *d_r, const uint8_t *d_s, const uint8_t *d_qx,
                            const uint8_t *d_qy, const uint8_t *d_dig, uint32_t dlen, uint8_t *d_ok, cudaStream_t st);
int sbv_launch_p384_w3_b64(sbv_engine *e, Dev &d, size_t n, const uint8_t *d_r, const uint8_t *d_s, const uint8_t *d_qx,
                            const uint8_t *d_qy, const uint8_t *d_dig, uint32_t dlen, uint8_t *d_ok, cudaStream_t st);
int sbv_init_gtables(sbv_engine *e, Dev &d);  // gtable.cu
int sbv_keys_build(sbv_engine *e, Dev &d);    // keyed.cu: (re)builds the per-key comb tables from the registry
void sbv_keys_free(Dev &d);
int sbv_launch_keyed(sbv_engine *e, Dev &d, uint8_t curve, size_t n, const uint32_t *d_slot, const uint8_t *d_r, const uint8_t *d_s,
                     const uint8_t *d_dig, uint32_t dlen, uint8_t *d_ok, cudaStream_t st);
int sbv_ensure_workspace(sbv_engine *e, Dev &d, size_t n);
int sbv_ensure_pinned(sbv_engine *e, Dev &d, size_t bytes);
int sbv_h2d(sbv_engine *e, Dev &d, void *dst, const void *src, size_t bytes, size_t &stage_off, cudaStream_t st);
int sbv_lane_acquire(sbv_engine *e);            // blocks until a lane index is free; returns it
void sbv_lane_release(sbv_engine *e, int lane);
int sbv_lane_ensure(sbv_engine *e, Dev &d, Dev::Lane &ln, size_t n, size_t pinned_bytes);
int sbv_lane_ensure_msgs(sbv_engine *e, Dev::Lane &ln, size_t bytes, size_t n_off);
// d_perm: n + 3072 words of scratch (may be null: no length sort)
int sbv_launch_sha256(sbv_engine *e, size_t n, const uint8_t *d_msgs, const uint64_t *d_off, uint64_t base, uint8_t *d_digest, uint32_t *d_perm,
                      cudaStream_t st);
int sbv_lane_h2d(sbv_engine *e, Dev::Lane &ln, void *dst, const void *src, size_t bytes, size_t &stage_off);
int sbv_ensure_scratch(sbv_engine *e, Dev &d, size_t bytes);
// takes the next scratch set of device d for a launch on stream st (waits for its previous user)
int sbv_take_scratch(sbv_engine *e, Dev &d, cudaStream_t st, Dev::Scratch **out);
