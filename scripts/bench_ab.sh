# usage: bash scripts/bench_ab.sh "ENV1=a ENV2=b" "ENV1=c" ...   -- short bench runs under different environments, alternating twice
for rep in 1 2; do
  for cfg in "$@"; do
    env $cfg timeout 300 python bench.py --steps 150 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$cfg', round(d['value'],1), round(d['e2e']['value'],1), round(d['e2e_dropin_loop']['value'],1), round(d['roofline']['launch_us'],1), round(d['cp_frame']['ms'],4))"
  done
done
