// per-key tables with 5-bit signed windows for P-256: construction + fixed-base verification
#include "inst_common.cuh"
using namespace sbv;
const KtOps sbv_kt5_p256 = {kt_geom<P256, 5>(), op_kt_build<P256, 5>, op_kt_verify<P256, 5>};
