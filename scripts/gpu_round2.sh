# Round evidence for the default (group) kernel: GPU tests, smoke, bench, ncu launch list + full capture of talker launches.
set -x
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/gpu_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e']['value'], d['e2e_dropin_loop']['value'], d['roofline']['launch_us'], round(d['roofline']['frac'],3), d['cp_frame']['ms'])"; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/plain_small.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_group.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qmk2_decode_kernel -s 70 -c 2 -o gpurun_out/prof_r01_group python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
