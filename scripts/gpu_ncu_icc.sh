set -x
timeout 600 python -m pytest tests -m gpu -q -x --timeout 400 > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gpu_tests.log
timeout 300 python scripts/perf_probe.py --configs "450,4500" > gpurun_out/probe.log 2>&1 && \
timeout 600 ncu --metrics sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gpu__time_duration.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio --clock-control none -k regex:qmk_decode_kernel -s 30 -c 2 --csv --log-file gpurun_out/icc.csv python scripts/perf_probe.py --configs "450,4500" > gpurun_out/ncu_icc.log 2>&1; echo "ncu rc=$?"; grep -E "poll_delay" gpurun_out/probe.log; python3 -c "
import csv
rows=[r for r in csv.reader(open('gpurun_out/icc.csv')) if len(r)>10]
for r in rows[1:6]: print(r[-3], r[-1])"
