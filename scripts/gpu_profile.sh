# ncu evidence for profiles/: launch list of the bench, then a full capture of one talker launch and one fused code-predictor frame.
set -x
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/plain_small.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qmk_decode_kernel -s 70 -c 2 -o gpurun_out/prof_r01_final python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
