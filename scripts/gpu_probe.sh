set -x
timeout 300 python scripts/perf_probe.py --configs "450,4500,0,1;450,4500,0,0;450,4500,0,1;450,4500,0,0" > gpurun_out/probe.log 2>&1; echo rc=$?; grep -E "poll_delay" gpurun_out/probe.log
