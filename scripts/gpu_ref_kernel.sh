#!/bin/bash
# try the upstream kernel (sm_100a build) a few times in fresh processes; each attempt is limited to 60 s
for i in 1 2 3; do
  timeout 70 python - <<'PY'
import sys, json
sys.path.insert(0, "."); import bench
print(json.dumps(bench.time_upstream_kernel_subprocess(45.0)))
PY
done
