set -x
timeout 500 python scripts/perf_probe.py --configs "800,4500,0,1,0,800,800;800,4500,0,1,0,400,800;800,4500,0,1,0,800,300;800,4500,0,1,0,400,300;800,4500,0,1,0,400,100;800,4500,0,1,0,200,100;800,4000,0,1,0,400,300;800,4500,0,1,0,800,800" > gpurun_out/probe.log 2>&1; echo rc=$?; grep -E "poll_delay" gpurun_out/probe.log
