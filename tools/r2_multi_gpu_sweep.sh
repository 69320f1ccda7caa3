#!/bin/bash
# Round-2 starter: validate and measure the fixed-capacity key exchange on N GPUs (default 2) in ONE gpurun call:
#   gpurun --gpus 2 --timeout 600 -- 'bash tools/r2_multi_gpu_sweep.sh 2'
# 1. NCCL parity of both exchange forms (tests/test_gpu_multi.py), 2. bench step time for the ragged exchange at 4
# sub-batches (the round-1 default: 24.1 ms on 2 GPUs) and for the fixed-capacity exchange at 4 / 8 / 16 sub-batches.
N=${1:-2}
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2
port=29540
run() {  # $1 = CTMR_FIXED_EXCHANGE, $2 = CTMR_BENCH_NSUB
  port=$((port + 1))
  CTMR_FIXED_EXCHANGE=$1 CTMR_BENCH_NSUB=$2 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" \
    --master-addr 127.0.0.1 --master-port $port bench.py --gpus "$N" --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/sweep_$1_$2.err |
    tail -1 > gpurun_out/sweep_$1_$2.json
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/sweep_%s_%s.json" % (sys.argv[1], sys.argv[2])))
    print("fixed=%s nsub=%s  %.1f M entries/s  %.2f ms/step  K_map %.2f ms" % (sys.argv[1], sys.argv[2], d["value"] / 1e6, d["ms_per_step"], d["roofline"]["kernel_ms_per_step"]))
except Exception as e:
    print("fixed=%s nsub=%s  FAILED: %r" % (sys.argv[1], sys.argv[2], e))
PY
}
mkdir -p gpurun_out
run 0 4
run 1 4
run 1 8
run 1 16
