#!/bin/bash
# one gpurun job: the latency probe (feature A/B grid) and the GPU parity tests most affected by the kernels it exercises
TAG=${1:-r02d}
mkdir -p gpurun_out
python tools/latency_probe.py > gpurun_out/${TAG}_latency.jsonl 2> gpurun_out/${TAG}_latency.err; echo "probe rc=$?"
python - <<PY
import json
for ln in open("gpurun_out/${TAG}_latency.jsonl"):
    r = json.loads(ln)
    print(r["D"], r["N"], "pdl", r["pdl"], "alt", r["altdir"], "fu", r["fused_update"], "ch", r["chunks"], "| cut wall", r["cut_wall_us"], "dev", r["dev_cut_us"],
          "prep", r["dev_prep_us"], "sweep", r["dev_sweep_us"], "merge", r["dev_merge_us"], "| omega", r["calc_omega_wall_us"], "upd", r["stochastic_updates_wall_us"], "tot", r["update_wall_us"],
          "bit", r["bit_identical_to_baseline"])
PY
tail -3 gpurun_out/${TAG}_latency.err
python -m pytest tests -m gpu -x -q -k "feasibility or parity or edges or golden or host_patch or end_to_end" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/${TAG}_pytest.log
