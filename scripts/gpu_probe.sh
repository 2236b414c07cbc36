set -x
timeout 400 python scripts/perf_probe.py --configs "800,4500" --trace > gpurun_out/probe.log 2>&1; echo rc=$?; tail -30 gpurun_out/probe.log
