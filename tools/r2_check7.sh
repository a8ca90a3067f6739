#!/bin/bash
# mma.sync linear backward: gradient parity + train step timing (f = FMA kernel, default = mma kernel)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_reference_trainer.py tests/test_gpu_parity.py -m gpu -q -x --timeout 300 2>&1 | tail -8 > gpurun_out/r2_bwd_tests.log
RGL_BWD_VARIANT=f timeout 300 python bench.py --workload train --batch 8192 --humans 10 --steps 40 --no-cpu-baseline > gpurun_out/r2_train_f.log 2>&1
timeout 300 python bench.py --workload train --batch 8192 --humans 10 --steps 40 --no-cpu-baseline > gpurun_out/r2_train_m.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 150 --csv --log-file gpurun_out/r2_launches_train.csv python bench.py --workload train --batch 8192 --humans 10 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_train_ncu.log 2>&1
cat gpurun_out/r2_bwd_tests.log
python - <<PY
import json
for f in ['gpurun_out/r2_train_f.log','gpurun_out/r2_train_m.log']:
    ok=False
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, 'ms/step %.4f'%d['ms_per_step'], 'samples/s %.4g'%d['value'], 'loss', d['final_loss'], 'launches', d['gpu_launches']); ok=True
    if not ok: print(f, open(f).read()[-1500:])
PY
