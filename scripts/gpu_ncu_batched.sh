set -x
timeout 400 python scripts/batched_probe.py > gpurun_out/batched_probe.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_tensor.sum,dram__bytes_read.sum --clock-control none -s 3000 -c 300 --csv --log-file gpurun_out/launches_batched.csv python scripts/batched_probe.py > gpurun_out/ncu_batched.log 2>&1; echo "ncu rc=$?"; cat gpurun_out/batched_probe.log
