#!/bin/bash
# 1 GPU: grouped ring, four observations per thread against two
TAG=${1:-r02l}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "grouped or trace or golden" 2>&1 | tail -5
for g4 in 1 0; do
SDGPU_GRP4=$g4 python tools/group_probe.py 4096x131072x2 4096x131072x3 4096x131072x4 4096x131072x8 8192x16384x2 16384x131072x2 > gpurun_out/${TAG}_group_probe_grp4_$g4.jsonl 2> gpurun_out/${TAG}_group_probe.err; echo "grp4=$g4 rc=$?"
python - <<PY
import json
for ln in open("gpurun_out/${TAG}_group_probe_grp4_$g4.jsonl"):
    r=json.loads(ln); print(r["lambda_rows"], r["bases"], r["observations"], "ldg", r["ldg_pairs_per_s"], "tma", r["tma_pairs_per_s"], "auto", r["auto_variant"], r["auto_pairs_per_s"], "grouped", r["grouped_pairs_per_s"], r["grouped_GBps_per_distinct_row"], r["identical"])
PY
done
tail -3 gpurun_out/${TAG}_group_probe.err
