set -x
timeout 300 python scripts/perf_probe.py --configs "450,4500,0;450,4500,1;300,4000,1;200,3500,1;600,5000,1" > gpurun_out/probe.log 2>&1; echo rc=$?; grep -E "poll_delay|repeated" gpurun_out/probe.log
timeout 300 python scripts/perf_probe.py --configs "450,4500,1" --trace > gpurun_out/probe_trace.log 2>&1; tail -19 gpurun_out/probe_trace.log
