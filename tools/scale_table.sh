#!/bin/bash
# Developer tool: 1 / 2 / 4 / 8-GPU table of bench.py on one multi-GPU box (profiles/r02_scale.md was made with it).
#   tools/scale_table.sh [cfg2] [cfg4]          (default: both)
# One line per (config, N): images/s, us per step, sharded_check.ok, e2e images/s; full JSON lines in gpurun_out/.
mkdir -p gpurun_out
cfgs=${@:-cfg2 cfg4}
port=29700
ngpu=$(python -c "import torch; print(torch.cuda.device_count())")
for cfg in $cfgs; do
  for n in 1 2 4 8; do
    [ "$n" -gt "$ngpu" ] && continue
    port=$((port + 1))
    out=gpurun_out/scale_${cfg}_n$n
    if [ "$n" = 1 ]; then
      python bench.py --config "$cfg" --no-cpu-baseline --no-other-configs > "$out.json" 2> "$out.err"
    else
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 \
        --master-port "$port" bench.py --gpus "$n" --config "$cfg" --no-cpu-baseline --no-other-configs > "$out.json" 2> "$out.err"
    fi
    python - "$out.json" "$cfg" "$n" <<'P'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[2], sys.argv[3], round(d["value"]), round(d["ms_per_step"] * 1e3, 1), d.get("sharded_check", {}).get("ok"),
          round(d["e2e"]["value"]))
except Exception as e:
    print(sys.argv[2], sys.argv[3], "FAILED", e)
P
  done
done
