#!/bin/bash
# one gpurun job (1 GPU): whole GPU test-suite incl. the long LP-driven runs, latency probe, FP64 peak, append benchmarks, bench line,
# ncu launch list + three `--set full` captures.  Output: gpurun_out/${TAG}_*
TAG=${1:-r02e}
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q --durations=12) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -22 gpurun_out/${TAG}_pytest_gpu.log
python tools/latency_probe.py > gpurun_out/${TAG}_latency.jsonl 2> gpurun_out/${TAG}_latency.err; echo "probe rc=$?"
python - <<PY
import json
for ln in open("gpurun_out/${TAG}_latency.jsonl"):
    r = json.loads(ln)
    print(r["D"], r["N"], "pdl", r["pdl"], "alt", r["altdir"], "fu", r["fused_update"], "| cut wall", r["cut_wall_us"], "dev", r["dev_cut_us"],
          "prep", r["dev_prep_us"], "sweep", r["dev_sweep_us"], "merge", r["dev_merge_us"], "| omega", r["calc_omega_wall_us"], "upd", r["stochastic_updates_wall_us"], "tot", r["update_wall_us"],
          "bit", r["bit_identical_to_baseline"])
PY
python tools/fp64_peak.py > gpurun_out/${TAG}_fp64_peak.jsonl 2> gpurun_out/${TAG}_fp64_peak.err; echo "fp64 rc=$?"; cut -c1-400 gpurun_out/${TAG}_fp64_peak.jsonl
python tools/rows_bench.py > gpurun_out/${TAG}_rows_bench.jsonl 2> gpurun_out/${TAG}_rows_bench.err; echo "rows rc=$?"; cut -c1-500 gpurun_out/${TAG}_rows_bench.jsonl
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/${TAG}_bench_n1.json
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "bench ref rc=$?"
# ncu: launch list of the bench (shares), then full captures of the three sweeps the verdict asked to re-capture
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_fullsize.csv \
	python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_sweep_tma -s 3 -c 1 -o gpurun_out/${TAG}_sweep_tma_full \
	python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_tma.log 2>&1; echo "ncu tma rc=$?"
SDGPU_SWEEP_VARIANT=1 ncu --set full --clock-control none --import-source on -k regex:k_sweep_ldg -s 3 -c 1 -o gpurun_out/${TAG}_sweep_ldg_full \
	python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_ldg.log 2>&1; echo "ncu ldg rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_sweep_tma_grp -s 2 -c 1 -o gpurun_out/${TAG}_sweep_grp \
	python tools/group_probe.py 4096x131072x4 > gpurun_out/${TAG}_ncu_grp.log 2>&1; echo "ncu grp rc=$?"
ls -la gpurun_out/${TAG}_*
