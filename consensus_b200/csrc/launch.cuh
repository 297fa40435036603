// launch.cuh — host launcher for one (curve, window, block) instantiation of the verify pipeline.
#pragma once
#include "engine.h"
#include "kernels.cuh"

namespace sbv {

template <class C, int W, int BLOCK, int MINB>
int launch_verify_t(sbv_engine *e, Dev &d, size_t n, const uint8_t *d_r, const uint8_t *d_s, const uint8_t *d_qx,
                    const uint8_t *d_qy, const uint8_t *d_dig, uint32_t dlen, uint8_t *d_ok, cudaStream_t st,
                    int curve_idx) {
    constexpr int S = 8;
    constexpr int TE = Windows<32 * C::N, W>::ENTRIES;
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
    CU(e, (launch_prep<C, W, S>(nn, d_r, d_s, d_dig, dlen, w->gidx, w->digits, w->flags, st)));
    if (ev) CU(e, cudaEventRecord(ev[1], st));
    const size_t smem = (size_t)TE * 3 * C::N * 4 * BLOCK;
    static bool attr_done = false;
    if (!attr_done) {
        CU(e, cudaFuncSetAttribute(k_verify<C, W, BLOCK, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    const uint32_t vblocks = (nn + BLOCK - 1) / BLOCK;
    k_verify<C, W, BLOCK, MINB><<<vblocks, BLOCK, smem, st>>>(nn, d_qx, d_qy, d_r, w->gidx, w->digits, w->flags,
                                                        reinterpret_cast<const uint4 *>(d.gtab[curve_idx]), d_ok);
    if (ev) CU(e, cudaEventRecord(ev[2], st));
    CU(e, cudaEventRecord(w->done, st));
    e->launches += 2;
    CU(e, cudaGetLastError());
    return 0;
}

// co-Z 4-bit-window variant (k_verify_coz)
template <class C, int BLOCK, int MINB, bool LOCKSTEP = false>
int launch_verify_coz_t(sbv_engine *e, Dev &d, size_t n, const uint8_t *d_r, const uint8_t *d_s, const uint8_t *d_qx,
                        const uint8_t *d_qy, const uint8_t *d_dig, uint32_t dlen, uint8_t *d_ok, cudaStream_t st, int curve_idx) {
    
    const uint32_t nn = (uint32_t)n;
    constexpr int S = 8;
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
    CU(e, (launch_prep<C, 4, S>(nn, d_r, d_s, d_dig, dlen, w->gidx, w->digits, w->flags, st)));
    if (ev) CU(e, cudaEventRecord(ev[1], st));
    const size_t smem = (size_t)7 * 2 * C::N * 4 * BLOCK;
    static bool attr_done = false;
    if (!attr_done) {
        CU(e, cudaFuncSetAttribute(k_verify_coz<C, BLOCK, MINB, LOCKSTEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    k_verify_coz<C, BLOCK, MINB, LOCKSTEP><<<(nn + BLOCK - 1) / BLOCK, BLOCK, smem, st>>>(
        nn, d_qx, d_qy, d_r, w->gidx, w->digits, w->flags, reinterpret_cast<const uint4 *>(d.gtab[curve_idx]), w->tscr, d_ok);
    if (ev) CU(e, cudaEventRecord(ev[2], st));
    CU(e, cudaEventRecord(w->done, st));
    e->launches += 2;
    CU(e, cudaGetLastError());
    return 0;
}

}  // namespace sbv

#define SBV_DEFINE_LAUNCHER_COZ_LOCKSTEP(NAME, CURVE, BLOCK, IDX)                                                            \
    int NAME(sbv_engine *e, Dev &d, size_t n, const uint8_t *d_r, const uint8_t *d_s, const uint8_t *d_qx,               \
             const uint8_t *d_qy, const uint8_t *d_dig, uint32_t dlen, uint8_t *d_ok, cudaStream_t st) {                 \
        return sbv::launch_verify_coz_t<sbv::CURVE, BLOCK, 1, true>(e, d, n, d_r, d_s, d_qx, d_qy, d_dig, dlen, d_ok, st, IDX); \
    }

#define SBV_DEFINE_LAUNCHER_COZ(NAME, CURVE, BLOCK, MINB, IDX)                                                              \
    int NAME(sbv_engine *e, Dev &d, size_t n, const uint8_t *d_r, const uint8_t *d_s, const uint8_t *d_qx,               \
             const uint8_t *d_qy, const uint8_t *d_dig, uint32_t dlen, uint8_t *d_ok, cudaStream_t st) {                 \
        return sbv::launch_verify_coz_t<sbv::CURVE, BLOCK, MINB>(e, d, n, d_r, d_s, d_qx, d_qy, d_dig, dlen, d_ok, st, IDX); \
    }

#define SBV_DEFINE_LAUNCHER(NAME, CURVE, W, BLOCK, MINB, IDX)                                                                 \
    int NAME(sbv_engine *e, Dev &d, size_t n, const uint8_t *d_r, const uint8_t *d_s, const uint8_t *d_qx,               \
             const uint8_t *d_qy, const uint8_t *d_dig, uint32_t dlen, uint8_t *d_ok, cudaStream_t st) {                 \
        return sbv::launch_verify_t<sbv::CURVE, W, BLOCK, MINB>(e, d, n, d_r, d_s, d_qx, d_qy, d_dig, dlen, d_ok, st, IDX); \
    }
