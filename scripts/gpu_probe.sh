set -x
timeout 600 python scripts/perf_probe.py --configs "1,0;8,0;8,1;16,0;16,1;4,0" > gpurun_out/probe_cfgs.log 2>&1; echo rc=$?
timeout 300 python scripts/perf_probe.py --configs "16,0" --trace > gpurun_out/probe_trace_r16.log 2>&1; echo rc=$?
cat gpurun_out/probe_cfgs.log; tail -20 gpurun_out/probe_trace_r16.log
