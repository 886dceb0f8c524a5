#!/bin/bash
# final measurements, call B: one ncu --set full capture per remaining configuration, summarised ON THE BOX (the reports are ~18 MB each)
mkdir -p gpurun_out
cap() { tag=$1; shift
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rx -s 1 -c 1 -f -o /tmp/$tag python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-configs "$@" > gpurun_out/$tag.log 2>&1
  tools/ncu_summary.sh /tmp/$tag.ncu-rep /tmp/$tag > /dev/null 2>&1
  cp /tmp/$tag.metrics.txt gpurun_out/${tag}_metrics.txt
  python tools/ncu_hotspots.py /tmp/$tag.source.csv 40 > gpurun_out/${tag}_source_hotspots.txt
  grep -E "Kernel Name|time_duration|dram__bytes|issue_active|inst_executed.sum" gpurun_out/${tag}_metrics.txt | cut -c1-160
}
cap r2_k_rx_ncu_bell103 --mode 300 --streams 32768 --amplitude 0.5
cap r2_k_rx_ncu_rtty --mode rtty --rate 8000 --streams 262144 --nsamples 32000
cap r2_k_rx_ncu_same --mode same --streams 131072 --nsamples 24000
cap r2_k_rx_ncu_cfg2_awgn --awgn 0.35
