#!/bin/bash
TAG=${1:-r02q}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "grouped" 2>&1 | tail -3
python tools/group_probe.py 4096x131072x2 4096x131072x3 4096x131072x4 4096x131072x8 8192x16384x2 16384x131072x2 > gpurun_out/${TAG}_group_probe.jsonl 2> gpurun_out/${TAG}_group_probe.err; echo "rc=$?"
python - <<PY
import json
for ln in open("gpurun_out/${TAG}_group_probe.jsonl"):
    r=json.loads(ln); print(r["lambda_rows"], r["bases"], r["observations"], "ldg", r["ldg_pairs_per_s"], "tma", r["tma_pairs_per_s"], "auto", r["auto_variant"], r["auto_pairs_per_s"], "grouped", r["grouped_pairs_per_s"], r["grouped_GBps_per_distinct_row"], r["identical"])
PY
tail -3 gpurun_out/${TAG}_group_probe.err
ncu --set full --clock-control none --import-source on -k regex:k_sweep_tma_grp4 -s 2 -c 1 -o gpurun_out/${TAG}_sweep_grp4 python tools/group_probe.py 4096x131072x4 > gpurun_out/${TAG}_ncu_grp4.log 2>&1; echo ncu rc=$?
