// engine.h — shared host-side state of libsbv.so (one engine = 1..8 devices of one box).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdint>
#include <condition_variable>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/sbv.h"
#include "ops.h"

constexpr int SBV_LANES = 6;    // concurrent host-buffer calls per engine
constexpr int SBV_SCRATCH = 8;  // scratch sets per device (> SBV_LANES + 1: a launch may be held open per lane)
constexpr int SBV_MAX_CHUNKS = 32;  // a large host-buffer batch is uploaded and verified in at most this many chunks

struct Dev {
    int ordinal = 0;
    cudaStream_t stream = nullptr;
    uint32_t *gtab[2] = {nullptr, nullptr};
    // Per-launch workspace of the verify pipeline.  A launch takes the next set and first waits for the event of
    // that set's previous user, so launches on different streams overlap without sharing mutable state.
    struct Scratch {
        size_t cap = 0;   // items
        size_t kcap = 0;  // keys with a table per launch
        uint32_t *uw = nullptr;     // k_prep output: u1, u2 word-major [2N][cap]
        uint8_t *flags = nullptr;   // r, s range verdicts
        uint32_t *tscr = nullptr;   // k_verify_coz per-signature scratch: 12N words, word-major
        uint32_t *gacc = nullptr;   // k_gpart -> k_verify_kt: u1*G of every item (Jacobian, 3N words, word-major)
        // key grouping
        uint32_t hsize = 0;
        uint32_t *htab = nullptr, *rep = nullptr, *keylist = nullptr, *klist = nullptr, *glist = nullptr;
        uint32_t *zeroed = nullptr;  // one memset: counters[4], kcnt[n], then counters[4] per chunk of a chunked launch
        int32_t *keyid = nullptr, *item_kid = nullptr;
        // per-key tables of the launch
        uint32_t *bases = nullptr, *hs = nullptr, *ztop = nullptr, *pref = nullptr, *ktab = nullptr;
        uint8_t *keyflags = nullptr;
        cudaStream_t s_tab = nullptr, s_gen = nullptr;  // table construction / generic kernel run beside the main stream
        cudaEvent_t done = nullptr, ev_group = nullptr, ev_prep = nullptr, ev_tab = nullptr, ev_gen = nullptr;
        bool used = false;
        bool open = false;  // taken by a launch whose second half has not been enqueued yet
        struct Caps {  // bytes allocated per buffer
            size_t uw = 0, flags = 0, tscr = 0, gacc = 0, htab = 0, rep = 0, keylist = 0, klist = 0, glist = 0, zeroed = 0, keyid = 0, item_kid = 0, bases = 0,
                   hs = 0, ztop = 0, pref = 0, ktab = 0, keyflags = 0;
        } caps;
    } ws[SBV_SCRATCH];
    unsigned ws_next = 0;
    // Per-call lanes of the host-buffer entry points: own stream, input/verdict buffers and pinned staging, so that
    // SBV_LANES host threads can have a call in flight each (H2D / kernels / D2H of one overlap the others').
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
        // small device scratch (quorum inputs / counts, packed verdict masks) and its pinned mirror
        uint8_t *d_aux = nullptr, *h_aux = nullptr;
        size_t aux_cap = 0;
        // second stream of the call (mixed-curve batches run their two pipelines side by side)
        cudaStream_t stream2 = nullptr;
        cudaEvent_t ev_a = nullptr, ev_b = nullptr;
        cudaEvent_t ev_chunk[SBV_MAX_CHUNKS] = {};  // "chunk c has arrived" (recorded on stream2, the upload stream of a chunked call)
    } lanes[SBV_LANES];
    // generic scratch of the entry points that serialise on the engine lock
    uint8_t *d_scratch = nullptr;
    size_t scratch_cap = 0;
    // registered keys (sbv_set_keys): per-curve tables (8-bit signed windows), validity flags, slot -> table index
    uint32_t *ktab[2] = {nullptr, nullptr};
    uint8_t *keyflags[2] = {nullptr, nullptr};
    int32_t *slot2local[2] = {nullptr, nullptr};
    uint32_t n_slots = 0, n_local[2] = {0, 0};
    // profiling: event quadruples per verify launch (start, after prep, before / after the dominant kernel)
    std::vector<cudaEvent_t> prof_events;
    size_t prof_used = 0;
};

struct sbv_engine {
    std::vector<Dev> devs;
    std::mutex mu;       // kernel enqueue + workspace growth
    std::mutex err_mu;   // last-error string
    std::condition_variable lane_cv;
    bool lane_busy[SBV_LANES] = {};
    std::string err;
    std::atomic<uint64_t> launches{0};
    int keyed_warp_limit = 2048;   // registered-key batches up to this size use one warp per signature (SBV_KEYED_WARP_LIMIT)
    int group_threshold = 16;      // a key gets a table when it occurs at least this often in a batch (SBV_GROUP_THRESHOLD; 0 = never)
    int group_max_keys = 8192;     // table slots per launch (SBV_GROUP_MAX_KEYS)
    int group_min_batch = 0;       // launches smaller than this skip the grouping (SBV_GROUP_MIN_BATCH)
    int chunk_items = 262144;      // host-buffer shards of >= this many items are uploaded and verified in >= 2 chunks of nominally this size (SBV_CHUNK_ITEMS; 0 = never)
    bool gsplit = true;            // u1*G in its own kernel beside the table construction (SBV_GSPLIT=0: inside the fixed-base kernel)
    uint32_t hash_seed = 0x9e3779b9u;
    bool profiling = false;
    // NCCL (loaded lazily with dlopen so single-device, single-rank users never touch it)
    void *nccl_lib = nullptr;
    std::vector<void *> nccl_comms;  // one per device of a multi-device engine
    // one-process-per-GPU deployments: this engine is rank `rank` of `nranks`.  One communicator per CHANNEL: the
    // collectives of a channel must be issued in the same order on every rank, so concurrent host threads take one each.
    std::vector<void *> rank_comms;
    // per channel: a high-priority stream for the pack + all-gather of a step (fork / join with two events), so that the
    // exchange is dispatched ahead of the pending blocks of other lanes' verification kernels (SBV_GATHER_PRIORITY)
    struct ChannelHi { cudaStream_t st = nullptr; cudaEvent_t in = nullptr, out = nullptr; };
    std::vector<ChannelHi> rank_hi;
    bool gather_hi = true;
    bool tab_hi = true;            // table-construction side streams at high priority: their few, latency-bound blocks are dispatched ahead of
                                   // the pending blocks of other launches' verification kernels (SBV_TAB_PRIORITY=0: e2e 72.9 -> 76 M/s with it)
    int rank = 0, nranks = 1;
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
    if (e) {
        std::lock_guard<std::mutex> lk(e->err_mu);
        e->err = buf;
    }
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

// one keys-per-item launch between its two halves (pipeline.cu)
struct VerifyLaunch {
    Dev::Scratch *w = nullptr;
    cudaEvent_t *ev = nullptr;
    const uint8_t *d_qx = nullptr, *d_qy = nullptr;
    size_t n = 0;
    uint8_t curve = 0;
    bool grouping = false;
    int chunks = 1;   // > 1: the second half comes chunk by chunk (sbv_launch_verify_chunk)
};

// ---- pipeline.cu: the verify pipelines (device pointers in, verdict bytes out; enqueue only, no sync) ----
// keys-per-item: k_prep, key grouping, per-key tables for repeated keys, fixed-base kernel + generic kernel for the rest
int sbv_launch_verify(sbv_engine *e, Dev &d, uint8_t curve, size_t n, const uint8_t *d_r, const uint8_t *d_s, const uint8_t *d_qx,
                      const uint8_t *d_qy, const uint8_t *d_dig, uint32_t dlen, uint8_t *d_ok, cudaStream_t st);
int sbv_launch_verify_begin(sbv_engine *e, Dev &d, uint8_t curve, size_t n, const uint8_t *d_qx, const uint8_t *d_qy, cudaStream_t st, VerifyLaunch *vl,
                            int chunks = 1);
void sbv_launch_verify_abort(const VerifyLaunch &vl, cudaStream_t st);  // hand the scratch set back after a fault between the halves
// second half for items [lo, lo + cn) of a launch begun with chunks > 1 (the pointers are those of the WHOLE batch);
// `last` closes the launch
int sbv_launch_verify_chunk(sbv_engine *e, Dev &d, const VerifyLaunch &vl, int c, size_t lo, size_t cn, bool last, const uint8_t *d_r, const uint8_t *d_s,
                            const uint8_t *d_dig, uint32_t dlen, uint8_t *d_ok, cudaStream_t st);
int sbv_launch_verify_finish(sbv_engine *e, Dev &d, const VerifyLaunch &vl, const uint8_t *d_r, const uint8_t *d_s, const uint8_t *d_dig, uint32_t dlen,
                             uint8_t *d_ok, cudaStream_t st);
// registered keys (sbv_set_keys)
int sbv_launch_keyed(sbv_engine *e, Dev &d, uint8_t curve, size_t n, const uint32_t *d_slot, const uint8_t *d_r, const uint8_t *d_s,
                     const uint8_t *d_dig, uint32_t dlen, uint8_t *d_ok, cudaStream_t st);
int sbv_init_gtables(sbv_engine *e, Dev &d);
int sbv_keys_build(sbv_engine *e, Dev &d);  // (re)builds the per-key tables of the registry
void sbv_keys_free(Dev &d);
void sbv_scratch_free(Dev &d);

// ---- engine.cu helpers shared with the other translation units ----
int sbv_lane_acquire(sbv_engine *e);            // blocks until a lane index is free; returns it
void sbv_lane_release(sbv_engine *e, int lane);
int sbv_lane_ensure(sbv_engine *e, Dev &d, Dev::Lane &ln, size_t n, size_t pinned_bytes);
int sbv_lane_ensure_msgs(sbv_engine *e, Dev::Lane &ln, size_t bytes, size_t n_off);
int sbv_lane_ensure_aux(sbv_engine *e, Dev::Lane &ln, size_t bytes);
// d_perm: n + 3072 words of scratch (may be null: no length sort)
int sbv_launch_sha256(sbv_engine *e, size_t n, const uint8_t *d_msgs, const uint64_t *d_off, uint64_t base, uint8_t *d_digest, uint32_t *d_perm,
                      cudaStream_t st);
int sbv_lane_h2d(sbv_engine *e, Dev::Lane &ln, void *dst, const void *src, size_t bytes, size_t &stage_off, cudaStream_t st = nullptr);
int sbv_ensure_scratch(sbv_engine *e, Dev &d, size_t bytes);

// A host-buffer call owns one lane on every device for its duration.  On every exit path — faults included — the
// lane's streams are drained before the lane is handed to the next caller, so no copy into the caller's buffers or out
// of the lane's staging area is still in flight when the call returns (cgo contract of sbv.h).
struct LaneGuard {
    sbv_engine *e;
    int lane;
    explicit LaneGuard(sbv_engine *eng) : e(eng), lane(sbv_lane_acquire(eng)) {}
    ~LaneGuard() {
        for (Dev &d : e->devs) {
            if (!d.lanes[lane].stream) continue;
            cudaSetDevice(d.ordinal);
            if (d.lanes[lane].stream2) cudaStreamSynchronize(d.lanes[lane].stream2);
            cudaStreamSynchronize(d.lanes[lane].stream);
        }
        sbv_lane_release(e, lane);
    }
    LaneGuard(const LaneGuard &) = delete;
    LaneGuard &operator=(const LaneGuard &) = delete;
};
