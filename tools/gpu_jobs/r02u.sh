#!/bin/bash
TAG=${1:-r02u}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_edges.py tests/test_host_patch.py -m gpu -q -x 2>&1 | tail -3
for sh in -1 0 1 2 3; do echo "shape $sh"; SDGPU_DC_SHAPE=$sh python tools/rows_bench.py 2>/dev/null | head -2 | python -c "
import sys, json
for ln in sys.stdin:
    r=json.loads(ln); print(r['duals'], r['observations'], 'new_obs', r['new_observation_us'], 'new_dual', r['new_dual_us'])
"; done
python tools/latency_probe.py --quick 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    r=json.loads(ln)
    if 'D' in r: print(r['D'], r['N'], 'pdl', r['pdl'], 'fu', r['fused_update'], 'cut', r['cut_wall_us'], 'omega', r['calc_omega_wall_us'], 'upd', r['stochastic_updates_wall_us'], 'tot', r['update_wall_us'], r['bit_identical_to_baseline'])
"
