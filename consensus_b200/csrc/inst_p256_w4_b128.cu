#include "launch.cuh"
SBV_DEFINE_LAUNCHER(sbv_launch_p256_w4_b128, P256, 4, 128, 2, 0)
