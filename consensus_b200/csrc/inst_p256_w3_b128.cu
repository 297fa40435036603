#include "launch.cuh"
SBV_DEFINE_LAUNCHER(sbv_launch_p256_w3_b128, P256, 3, 128, 3, 0)
