// gtable.cu — one-time construction of the fixed-base comb tables on the device.
#include "engine.h"
#include "kernels.cuh"

int sbv_init_gtables(sbv_engine *e, Dev &d) {
    CU(e, cudaMalloc(&d.gtab[0], (size_t)32 * 256 * 16 * 4));
    CU(e, cudaMalloc(&d.gtab[1], (size_t)48 * 256 * 24 * 4));
    sbv::k_gtable_init<sbv::P256><<<32 * 256 / 128, 128, 0, d.stream>>>(d.gtab[0]);
    sbv::k_gtable_init<sbv::P384><<<48 * 256 / 128, 128, 0, d.stream>>>(d.gtab[1]);
    e->launches += 2;
    CU(e, cudaGetLastError());
    CU(e, cudaStreamSynchronize(d.stream));
    return 0;
}
