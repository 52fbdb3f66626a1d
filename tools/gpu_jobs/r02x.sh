#!/bin/bash
# 1 GPU: parity + latency grid + full-size launch list after the occupancy fix of the merge kernel and the prologue
TAG=${1:-r02x}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_edges.py tests/test_sd_end_to_end.py tests/test_host_patch.py -m gpu -q -x 2>&1 | tail -3
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
python bench.py --steps 30 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
r=json.loads(open("gpurun_out/${TAG}_bench_n1.json").read().strip().splitlines()[-1])
print({k:r[k] for k in ("value","ms_per_step")}, r["step_split_ms_rank0"], r["e2e"]["ms_per_step"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_fullsize.csv \
	python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/${TAG}_launches_fullsize.csv")) if len(r) > 14 and r[0].isdigit()]
acc = collections.defaultdict(list)
for r in rows: acc[r[4].split("(")[0]].append(float(r[14]) / 1e3)
for k, v in acc.items(): print(f"{k}: n={len(v)} mean={sum(v)/len(v):.1f} us")
PY
