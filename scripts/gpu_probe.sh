set -x
timeout 300 python scripts/perf_probe.py --configs "450,4500" --trace > gpurun_out/probe_min.log 2>&1; echo rc=$?; tail -20 gpurun_out/probe_min.log
