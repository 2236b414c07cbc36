set -x
timeout 300 python scripts/perf_probe.py --configs "450,4500" > gpurun_out/plain_probe.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qmk_decode_kernel -s 30 -c 1 -o gpurun_out/prof_talker_v2 python scripts/perf_probe.py --configs "450,4500" > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_full.log
