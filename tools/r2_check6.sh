#!/bin/bash
# quad kernel (RGL_GRAPH_VARIANT=q): parity tests + timings against p
mkdir -p gpurun_out
RGL_GRAPH_VARIANT=q timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc_variants.py tests/test_planner.py -m gpu -q -x --timeout 300 2>&1 | tail -12 > gpurun_out/r2_tq_tests.log
for v in p q; do
  RGL_GRAPH_VARIANT=$v timeout 300 python tools/quick_time.py > gpurun_out/r2_qt2_$v.log 2>&1
  RGL_GRAPH_VARIANT=$v timeout 200 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_bench2_$v.log 2>&1
done
cat gpurun_out/r2_tq_tests.log
for v in p q; do echo "== variant $v"; cat gpurun_out/r2_qt2_$v.log; python - <<PY
import json
for l in open('gpurun_out/r2_bench2_$v.log'):
    if l.startswith('{'):
        d=json.loads(l); print('value %.4g'%d['value'], 'launch_us', d['roofline']['launch_us'], 'steady', d['extra']['steady_state']['value'], 'vp', d['extra']['value_path']['value'], 'e2e', d['e2e']['value'])
PY
done
