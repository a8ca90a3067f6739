#!/bin/bash
# PDL on/off A/B: tests, kernel timings, headline bench, planner
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -8 > gpurun_out/r2_gpu_tests_pdl.log
for pdl in 0 1; do
  RGL_PDL=$pdl QT_FAST=1 timeout 200 python tools/quick_time.py > gpurun_out/r2_qt_pdl$pdl.log 2>&1
  RGL_PDL=$pdl timeout 200 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_pdl$pdl.log 2>&1
  RGL_PDL=$pdl timeout 200 python bench.py --workload plan --no-cpu-baseline --steps 20 > gpurun_out/r2_plan_pdl$pdl.log 2>&1
  RGL_PDL=$pdl timeout 200 python bench.py --workload plan --no-cpu-baseline --steps 5 --humans 20 --depth 3 --roots 2048 --speed-samples 5 --rotation-samples 16 > gpurun_out/r2_plan5_pdl$pdl.log 2>&1
done
cat gpurun_out/r2_gpu_tests_pdl.log
for pdl in 0 1; do echo "== PDL $pdl"; cat gpurun_out/r2_qt_pdl$pdl.log; python - <<PY
import json
for f in ['gpurun_out/r2_bench_pdl$pdl.log','gpurun_out/r2_plan_pdl$pdl.log','gpurun_out/r2_plan5_pdl$pdl.log']:
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, 'value %.4g'%d['value'], 'ms/step %.5f'%d['ms_per_step'], 'launch_us', d.get('roofline',{}).get('launch_us'), 'e2e', d.get('e2e',{}).get('value'), 'vp', d.get('extra',{}).get('value_path'))
            break
    else:
        print(f, 'NO JSON'); print(open(f).read()[-1500:])
PY
done
