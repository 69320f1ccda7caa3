#!/bin/bash
# A/B of K_map variants on one B200 (run under gpurun).  Usage: tools/ab_map_variants.sh [entries] [workload]
N=${1:-4000000}
W=${2:-cfg2}
B="python bench.py --workload $W --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --entries $N"
P='import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print("RESULT %7.1f Mentries/s  step_ms %7.3f  map_ms/step %7.3f  sha_GB/s(map) %6.0f  hbm_frac %.4f" % (d["value"]/1e6, d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["config"]["bytes_per_gpu_per_step"]/d["roofline"]["kernel_ms_per_step"]/1e6, d["roofline"]["frac"]))'
while read -r v; do
  [ -z "$v" ] && continue
  echo "== $W $v"; env $v $B 2>&1 | python -c "$P" || env $v $B 2>&1 | tail -5
done <<'LIST'
CTMR_MAP_VARIANT=2
CTMR_MAP_VARIANT=2 CTMR_MAP_ROLLED=0
LIST
