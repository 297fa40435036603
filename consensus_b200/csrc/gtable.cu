// gtable.cu — one-time construction of the fixed-base comb tables on the device.
#include "engine.h"
#include "kernels.cuh"

int sbv_init_gtables(sbv_engine *e, Dev &d) {
    const size_t e256 = (size_t)sbv::P256::GWINS << sbv::P256::GW, e384 = (size_t)sbv::P384::GWINS << sbv::P384::GW;
    CU(e, cudaMalloc(&d.gtab[0], e256 * 16 * 4));  // 64 MiB: stays resident in the 126 MB L2
    CU(e, cudaMalloc(&d.gtab[1], e384 * 24 * 4));
    sbv::k_gtable_init<sbv::P256><<<(unsigned)((e256 + 127) / 128), 128, 0, d.stream>>>(d.gtab[0]);
    sbv::k_gtable_init<sbv::P384><<<(unsigned)((e384 + 127) / 128), 128, 0, d.stream>>>(d.gtab[1]);
    e->launches += 2;
    CU(e, cudaGetLastError());
    CU(e, cudaStreamSynchronize(d.stream));
    return 0;
}
