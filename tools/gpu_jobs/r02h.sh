#!/bin/bash
# 1 GPU: launch round-trip floor, append benchmarks after the delta-kernel revert, whole GPU suite, smoke, bench line (both arms), launch list
TAG=${1:-r02h}
mkdir -p gpurun_out
python tools/latency_probe.py --floor > gpurun_out/${TAG}_launch_floor.jsonl 2> gpurun_out/${TAG}_launch_floor.err; echo "floor rc=$?"; cat gpurun_out/${TAG}_launch_floor.jsonl
python tools/rows_bench.py > gpurun_out/${TAG}_rows_bench.jsonl 2> gpurun_out/${TAG}_rows_bench.err; echo "rows rc=$?"; cut -c1-600 gpurun_out/${TAG}_rows_bench.jsonl
(time python -m pytest tests -m gpu -q --durations=8) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -16 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"; cut -c1-1800 gpurun_out/${TAG}_bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_fullsize.csv \
	python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/${TAG}_launches_fullsize.csv")) if len(r) > 14 and r[0].isdigit()]
acc = collections.defaultdict(list)
for r in rows: acc[r[4].split("(")[0]].append(float(r[14]) / 1e3)
for k, v in acc.items(): print(f"{k}: n={len(v)} mean={sum(v)/len(v):.1f} us")
PY
