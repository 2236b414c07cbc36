#!/bin/bash
# round 2: bench (full config, driver-style short config, reference arm)
set -x
timeout 900 python bench.py --steps 500 --warmup 10 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_k20.json 2> gpurun_out/r2_bench_k20.err; echo "bench20 rc=$?"
tail -5 gpurun_out/r2_bench.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench.json", "gpurun_out/r2_bench_k20.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "async", round(d["e2e_per_frame_api"]["value"], 1),
              "dropin", round(d["e2e_dropin_loop"]["value"], 1), "upstream", round(d["e2e_upstream_loop"]["value"], 1), "reps", d["reps"])
        print("  roofline", round(d["roofline"]["frac"], 4), d["roofline"]["launch_us"], {k: (round(v["launch_us"], 1), round(v["frac"], 3)) for k, v in d["roofline_by_position"].items()})
        print("  cp", d["cp_frame"]["ms"], d["cp_frame"]["ms_sampled"], "batched", {k: round(v["ms_per_step"], 3) for k, v in d.get("batched", {}).items()}, d.get("cpu_baseline"))
    except Exception as e:
        print(f, "unreadable", e)
PY
