#!/bin/bash
# 1 GPU: parity subset + latency grid after the merge / prologue batching changes
TAG=${1:-r02i}
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_edges.py tests/test_sd_end_to_end.py tests/test_host_patch.py -m gpu -q -x) > gpurun_out/${TAG}_pytest_subset.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest_subset.log
python tools/latency_probe.py > gpurun_out/${TAG}_latency.jsonl 2> gpurun_out/${TAG}_latency.err; echo "probe rc=$?"
python - <<PY
import json
for ln in open("gpurun_out/${TAG}_latency.jsonl"):
    r = json.loads(ln)
    if "D" not in r: print(r); continue
    print(r["D"], r["N"], "pdl", r["pdl"], "alt", r["altdir"], "fu", r["fused_update"], "| cut wall", r["cut_wall_us"], "dev", r["dev_cut_us"],
          "prep", r["dev_prep_us"], "sweep", r["dev_sweep_us"], "merge", r["dev_merge_us"], "| omega", r["calc_omega_wall_us"], "upd", r["stochastic_updates_wall_us"], "tot", r["update_wall_us"],
          "bit", r["bit_identical_to_baseline"])
PY
