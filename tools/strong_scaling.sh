#!/bin/bash
# Strong-scaling record + multi-GPU parity on an 8-GPU box:
#   bench.py --scaling strong (fixed 8 192 x 1 048 576 x 256 table = 64 GiB, observations split N ways) at N = 1, 2, 4, 8, each run with the
#   sharded-vs-unsharded parity self-check and with the NCCL all-reduce and the NVLink peer exchange timed side by side;
#   tools/multigpu_check.py at world 2, 4, 8 and tools/group_check.py over all 8 GPUs.
# The small runs share the box (disjoint GPUs) to keep the 8x-charged wall time short; N = 8 runs alone.
# usage: bash tools/strong_scaling.sh <out-prefix>     e.g. gpurun_out/r02_strong
set -u
OUT=${1:-gpurun_out/r02_strong}
STEPS=${STEPS:-20}
run_bench() {   # <gpus csv> <n> <port>
	CUDA_VISIBLE_DEVICES=$1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $3 \
		bench.py --gpus $2 --scaling strong --steps $STEPS --warmup 3 --no-cpu > ${OUT}_n$2.json 2> ${OUT}_n$2.err
	echo "strong n=$2 rc=$?"
}
run_check() {   # <gpus csv> <n> <port>
	CUDA_VISIBLE_DEVICES=$1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $3 \
		tools/multigpu_check.py > ${OUT}_multigpu_check_w$2.log 2>&1
	echo "multigpu_check world=$2 rc=$? $(grep -h MULTIGPU_OK ${OUT}_multigpu_check_w$2.log)"
}
nvidia-smi -L > ${OUT}_gpus.txt
# phase 1: N = 4 on GPUs 0-3, N = 2 on 4-5, N = 1 on 6, parity check world 2 on ... (wait: GPUs are taken) -> checks in phase 2
run_bench 0,1,2,3 4 29601 &
run_bench 4,5 2 29602 &
CUDA_VISIBLE_DEVICES=6 python bench.py --gpus 1 --scaling strong --steps $STEPS --warmup 3 --no-cpu > ${OUT}_n1.json 2> ${OUT}_n1.err &
wait
echo "strong n=1 done"
# phase 2: the parity checks, world 2 and 4 side by side, then world 8 and the single-process group
run_check 0,1 2 29611 &
run_check 2,3,4,5 4 29612 &
wait
run_check 0,1,2,3,4,5,6,7 8 29613
python tools/group_check.py > ${OUT}_group_check_8gpu.log 2>&1; echo "group_check rc=$? $(grep -h GROUP_OK ${OUT}_group_check_8gpu.log)"
# phase 3: N = 8 alone
run_bench 0,1,2,3,4,5,6,7 8 29621
python - <<PY
import json, glob
rows = {}
for n in (1, 2, 4, 8):
    try:
        rows[n] = json.loads(open("${OUT}_n%d.json" % n).read().strip().splitlines()[-1])
    except Exception as exc:
        print("n=%d: no line (%s)" % (n, exc))
if 1 in rows:
    base = rows[1]["ms_per_step"]
    for n, r in sorted(rows.items()):
        col = r.get("collectives", {})
        print("N=%d ms/step %.4f speed-up %.3f efficiency %.3f split %s parity %s collectives %s" % (n, r["ms_per_step"], base / r["ms_per_step"], base / r["ms_per_step"] / n,
              {k: round(v, 4) for k, v in r["step_split_ms_rank0"].items()}, r.get("multi_gpu_parity", {}).get("nccl", "-") + "/" + r.get("multi_gpu_parity", {}).get("peer", "-"),
              {k: round(v["ms_per_step"], 4) for k, v in col.items()}))
PY
