#!/bin/bash
# round 2 ncu evidence: launch list of the bench command, full captures of a talker launch, a code-predictor frame and the
# frame-loop launch, tensor-pipe utilisation of the batched kernels.  Every command first runs WITHOUT ncu.
set -x
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_plain_small.json 2> gpurun_out/r2_plain_small.err; echo "plain rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --min-seconds 0.01 > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 300 python scripts/ncu_case.py --batched > gpurun_out/r2_ncu_case_plain.log 2>&1; echo "case rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:qmk2_decode_kernel -s 8 -c 6 -o gpurun_out/prof_r02_b1 python scripts/ncu_case.py > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/r2_ncu_full.log
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,dram__bytes_read.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"qmk_bgemm_kernel|qmk_bstep_kernel|kb_" -c 700 --csv --log-file gpurun_out/r2_batched_tensor.csv python scripts/ncu_case.py --batched > gpurun_out/r2_ncu_batched.log 2>&1; echo "ncu batched rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_*.csv
