#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^(k_verify_keyed|k_prep)$" -s 2 -c 2 -f -o gpurun_out/prof_keyed python tools/prof_keyed.py > gpurun_out/ncu_keyed.log 2>&1
tail -2 gpurun_out/ncu_keyed.log; ls -la gpurun_out/prof_keyed.ncu-rep
