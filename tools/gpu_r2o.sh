#!/bin/bash
mkdir -p gpurun_out
cd experiments/tcgen05_blocksum
timeout 120 ./tc_blocksum 8 > ../../gpurun_out/r2o_tc.json 2> ../../gpurun_out/r2o_tc.err; echo rc=$?
cat ../../gpurun_out/r2o_tc.json; tail -3 ../../gpurun_out/r2o_tc.err
timeout 300 ncu --set full --clock-control none -k regex:k_blocksum -c 2 -f -o ../../gpurun_out/r2o_tc ./tc_blocksum 2 > ../../gpurun_out/r2o_ncu.log 2>&1
tail -2 ../../gpurun_out/r2o_ncu.log
