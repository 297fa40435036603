// keys-per-item kernel (k_verify_coz) for P-256
#include "inst_common.cuh"
using namespace sbv;
cudaError_t sbv_coz_p256(uint32_t n, const uint8_t *qx, const uint8_t *qy, const uint8_t *r, const uint32_t *uw, const uint8_t *flags,
                       const uint32_t *gtab, uint32_t *tscr, uint8_t *ok, const uint32_t *list, const uint32_t *count, cudaStream_t st) {
    return op_coz<P256>(n, qx, qy, r, uw, flags, gtab, tscr, ok, list, count, st);
}
