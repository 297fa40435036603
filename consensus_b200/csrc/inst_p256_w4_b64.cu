#include "launch.cuh"
SBV_DEFINE_LAUNCHER(sbv_launch_p256_w4_b64, P256, 4, 64, 0)
