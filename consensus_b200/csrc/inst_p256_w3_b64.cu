#include "launch.cuh"
SBV_DEFINE_LAUNCHER(sbv_launch_p256_w3_b64, P256, 3, 64, 7, 0)
