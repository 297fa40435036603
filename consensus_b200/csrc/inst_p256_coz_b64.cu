#include "launch.cuh"
SBV_DEFINE_LAUNCHER_COZ(sbv_launch_p256_coz_b64, P256, 64, 7, 0)
