#include "launch.cuh"
SBV_DEFINE_LAUNCHER(sbv_launch_p384_w3_b64, P384, 3, 64, 4, 1)
