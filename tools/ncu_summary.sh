#!/bin/bash
# tools/ncu_summary.sh <report.ncu-rep> [out-prefix]: the metrics and the per-source-line hot spots the profiles/ files hold
rep=$1; out=${2:-/tmp/ncu}
ncu -i $rep --page raw --csv 2>/dev/null | python3 -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]; units=rows[1]; vals=rows[2]
keep=['dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','gpu__time_duration.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','launch__block_size','launch__grid_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__registers_per_thread','launch__shared_mem_per_block_dynamic','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__warps_active.avg.per_cycle_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__warps_eligible.avg.per_cycle_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','lts__t_bytes.sum','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','smsp__inst_executed_op_local_ld.sum','smsp__inst_executed_op_local_st.sum']
for i,h in enumerate(hdr):
    if h in keep or h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio') or h=='Kernel Name':
        print('%-100s %s %s'%(h, vals[i], units[i]))
" > $out.metrics.txt
ncu -i $rep --page source --csv --print-source cuda,sass 2>/dev/null > $out.source.csv
python3 - $out.source.csv > $out.hotspots.txt <<'PY'
import csv,sys,collections
rows=list(csv.reader(open(sys.argv[1])))
# find header
hi=[i for i,r in enumerate(rows) if 'Source' in r and ('# Instructions Executed' in ' '.join(r) or 'Instructions Executed' in ' '.join(r))]
print('rows',len(rows),'hdr at',hi[:3])
if hi:
    h=rows[hi[0]]
    def col(name):
        for i,c in enumerate(h):
            if c.strip()==name: return i
        return None
    ci=col('Instructions Executed'); cs=col('Warp Stall Sampling (All Samples)') or col('Warp Stall Sampling (All Cycles)'); src=col('Source')
    print(h)
PY
head -60 $out.metrics.txt
