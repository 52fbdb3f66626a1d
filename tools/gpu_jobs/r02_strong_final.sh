#!/bin/bash
# 8-GPU box: strong-scaling record at HEAD (fixed 8 192 x 1 048 576 table): N = 4, 2, 1 side by side on disjoint GPUs, then N = 8 with both collectives
OUT=${1:-gpurun_out/r02_strong_final}
STEPS=${STEPS:-20}
run_bench() {   # <gpus csv> <n> <port> [extra]
	CUDA_VISIBLE_DEVICES=$1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $3 \
		bench.py --gpus $2 --scaling strong --steps $STEPS --warmup 3 --no-cpu $4 > ${OUT}_n$2.json 2> ${OUT}_n$2.err
	echo "strong n=$2 rc=$?"
}
run_bench 0,1,2,3 4 29601 "" &
run_bench 4,5 2 29602 "" &
CUDA_VISIBLE_DEVICES=6 python bench.py --gpus 1 --scaling strong --steps $STEPS --warmup 3 --no-cpu > ${OUT}_n1.json 2> ${OUT}_n1.err &
wait
run_bench 0,1,2,3,4,5,6,7 8 29603 "--both-collectives"
python - <<PY
import json
base=None
for n in (1,2,4,8):
    try: r=json.loads(open("${OUT}_n%d.json" % n).read().strip().splitlines()[-1])
    except Exception as e: print(n, "failed", e); continue
    if n==1: base=r["ms_per_step"]
    print(n, "ms", round(r["ms_per_step"],4), "eff", round(base/(n*r["ms_per_step"]),4) if base else None, r.get("config",{}).get("collective"), r.get("multi_gpu_parity"), {k:round(v["ms_per_step"],4) for k,v in r.get("collectives",{}).items()}, r.get("step_split_ms_rank0"))
PY
