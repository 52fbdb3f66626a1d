#!/bin/bash
TAG=${1:-r02o}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "grouped" 2>&1 | tail -3
for cfg in "1 0.5" "1 0.05" "1 0" "0 0.5"; do set -- $cfg
SDGPU_SEED=$1 PROBE_FRESH_X=$2 python tools/group_probe.py 4096x131072x2 4096x131072x4 4096x131072x8 16384x131072x2 > gpurun_out/${TAG}_group_probe_seed$1_x$2.jsonl 2> gpurun_out/${TAG}_group_probe.err; echo "seed=$1 x=$2 rc=$?"
python - <<PY
import json
for ln in open("gpurun_out/${TAG}_group_probe_seed$1_x$2.jsonl"):
    r=json.loads(ln); print(r["lambda_rows"], r["bases"], r["observations"], "tma", r["tma_pairs_per_s"], "auto", r["auto_variant"], r["auto_pairs_per_s"], "grouped", r["grouped_pairs_per_s"], r["grouped_GBps_per_distinct_row"], r["identical"])
PY
done
tail -3 gpurun_out/${TAG}_group_probe.err
