#!/bin/bash
mkdir -p gpurun_out
FSK_B200_MULTI=2 bash tools/gpu_ncu.sh r2h_bell103_m2 --mode 300 --streams 8192 --amplitude 0.5
FSK_B200_MULTI=2 bash tools/gpu_ncu.sh r2h_rtty_m2 --mode rtty --rate 8000 --streams 65536 --nsamples 32000
FSK_B200_MULTI=0 bash tools/gpu_ncu.sh r2h_cfg2_awgn_m0 --streams 16384 --awgn 0.35
du -sh gpurun_out
