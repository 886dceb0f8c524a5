#!/bin/bash
# round-1 closing run on the GPU box: parity, smoke, bench, ncu evidence (outputs under gpurun_out/)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -2 > gpurun_out/pytest_final.log; cat gpurun_out/pytest_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final.log 2>&1; tail -1 gpurun_out/smoke_final.log
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_rx -c 1 -f -o gpurun_out/prof_r1_final2 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_final2.log 2>&1; echo ncu-full done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/launches_final.log 2>&1; echo ncu-launches done
timeout 400 python bench.py --impl reference > gpurun_out/bench_reference_final.json 2> gpurun_out/bench_reference_final.err; tail -c 400 gpurun_out/bench_reference_final.json
