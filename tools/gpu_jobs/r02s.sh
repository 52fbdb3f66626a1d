#!/bin/bash
TAG=${1:-r02s}
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q --durations=5) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/${TAG}_pytest_gpu.log
python tools/group_probe.py 4096x131072x1 4096x131072x2 4096x131072x3 4096x131072x4 4096x131072x8 8192x16384x2 16384x131072x2 > gpurun_out/${TAG}_group_probe.jsonl 2> gpurun_out/${TAG}_group_probe.err; echo "rc=$?"
python - <<PY
import json
for ln in open("gpurun_out/${TAG}_group_probe.jsonl"):
    r=json.loads(ln); print(r["lambda_rows"], r["bases"], r["observations"], "ldg", r["ldg_pairs_per_s"], "tma", r["tma_pairs_per_s"], "auto", r["auto_variant"], r["auto_pairs_per_s"], "grouped", r["grouped_pairs_per_s"], r["grouped_GBps_per_distinct_row"], r["identical"])
PY
ncu --set full --clock-control none --import-source on -k regex:k_sweep_tma_grp -s 2 -c 1 -o gpurun_out/${TAG}_sweep_grp python tools/group_probe.py 4096x131072x4 > gpurun_out/${TAG}_ncu_grp.log 2>&1; echo ncu rc=$?
