#!/bin/bash
# the rx parity tests with the prefix-table search forced wherever it fits, then the quick bench
mkdir -p gpurun_out
( time FSK_B200_PREFIX=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants or baseline_configs or reference_vectors or noise_sweep or s16 or chunks or lane_split or edge_cases or overflow or roundtrip or live_receiver" ) > gpurun_out/r2p_pytest.log 2>&1; tail -4 gpurun_out/r2p_pytest.log
bash tools/gpu_r2p_quick.sh "$@"
