#include "launch.cuh"
SBV_DEFINE_LAUNCHER_COZ_LOCKSTEP(sbv_launch_p256_coz_b448, P256, 448, 0)
