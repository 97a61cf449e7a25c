#!/bin/bash
# BASELINE configs[4]: optimiser sweep at 2M items - B in {4k .. 64k} x {hybrid AdamW+SparseAdam, all-sparse SparseAdam,
# all-dense AdamW semantics}.  One bench.py line per point, appended to $OUT (default gpurun_out/r2_config5_sweep.jsonl).
#   bash scripts/sweep_config5.sh [N_GPUS]
N=${1:-1}
OUT=${OUT:-gpurun_out/r2_config5_sweep_n$N.jsonl}
: > "$OUT"
for mode in hybrid sparse dense; do
  for B in 4096 8192 16384 32768 65536; do
    if [ "$N" = "1" ]; then
      python bench.py --mode $mode --batch $B --steps 20 --warmup 4 --no-cpu-baseline --no-retrieval --no-hook --no-fp32 >> "$OUT" 2>> "${OUT%.jsonl}.err" || echo "{\"error\": \"mode=$mode B=$B\"}" >> "$OUT"
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --mode $mode --batch $B --steps 20 --warmup 4 --no-cpu-baseline --no-retrieval >> "$OUT" 2>> "${OUT%.jsonl}.err" || echo "{\"error\": \"mode=$mode B=$B\"}" >> "$OUT"
    fi
  done
done
python - "$OUT" <<'PY'
import json, sys
for ln in open(sys.argv[1]):
    try:
        d = json.loads(ln)
        print(f'{d["config"]["workload"][:40]}... n={d["n_gpus"]} ms/step {d["ms_per_step"]:.3f} samples/s {d["value"]/1e6:.2f} M  e2e {d["e2e"]["value"]/1e6:.2f} M  | {d["config"]["workload"].split("B=")[1][:60]}')
    except Exception as e:
        print("ERR", ln[:100])
PY
