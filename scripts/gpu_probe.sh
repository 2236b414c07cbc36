set -x
timeout 500 python scripts/perf_probe.py --configs "800,4500,0,1,0;800,4500,0,1,1;800,3500,0,1,1;800,2500,0,1,1;800,1500,0,1,1;800,4500,0,1,0;800,3500,0,1,1" > gpurun_out/probe.log 2>&1; echo rc=$?; grep -E "poll_delay|repeated" gpurun_out/probe.log
