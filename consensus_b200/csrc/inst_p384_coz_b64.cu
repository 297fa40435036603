#include "launch.cuh"
SBV_DEFINE_LAUNCHER_COZ(sbv_launch_p384_coz_b64, P384, 64, 4, 1)
