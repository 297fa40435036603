#include "launch.cuh"
SBV_DEFINE_LAUNCHER(sbv_launch_p384_w3_b128, P384, 3, 128, 1, 1)
