#!/bin/bash
# 1 GPU: basis-chunk count at the strong-scaling per-GPU shape, latency probe (prologue reverted), ssn iterations/s at k = 2 500
TAG=${1:-r02f}
mkdir -p gpurun_out
: > gpurun_out/${TAG}_chunk_probe.jsonl
for c in 0 7 10 14 19 28 42 55; do SDGPU_CHUNKS=$c python tools/chunk_probe.py 8192 131072 >> gpurun_out/${TAG}_chunk_probe.jsonl 2>> gpurun_out/${TAG}_chunk_probe.err; done
for c in 0 4 7 14; do SDGPU_CHUNKS=$c python tools/chunk_probe.py 8192 1048576 >> gpurun_out/${TAG}_chunk_probe.jsonl 2>> gpurun_out/${TAG}_chunk_probe.err; done
cat gpurun_out/${TAG}_chunk_probe.jsonl
python tools/latency_probe.py > gpurun_out/${TAG}_latency.jsonl 2> gpurun_out/${TAG}_latency.err; echo "probe rc=$?"
python - <<PY
import json
for ln in open("gpurun_out/${TAG}_latency.jsonl"):
    r = json.loads(ln)
    print(r["D"], r["N"], "pdl", r["pdl"], "alt", r["altdir"], "fu", r["fused_update"], "| cut wall", r["cut_wall_us"], "dev", r["dev_cut_us"],
          "prep", r["dev_prep_us"], "sweep", r["dev_sweep_us"], "merge", r["dev_merge_us"], "| omega", r["calc_omega_wall_us"], "upd", r["stochastic_updates_wall_us"], "tot", r["update_wall_us"],
          "bit", r["bit_identical_to_baseline"])
PY
python tools/sd_iterations_bench.py --iterations 2500 --backends gpu,reference > gpurun_out/${TAG}_sd_iterations_ssn_k2500.jsonl 2> gpurun_out/${TAG}_sd_iterations.err; echo "sd iterations rc=$?"
cut -c1-700 gpurun_out/${TAG}_sd_iterations_ssn_k2500.jsonl
