#!/bin/bash
# full round check on 2 GPUs: GPU tests, smoke, default bench at N=1 (both arms) and N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -6 > gpurun_out/r2_gpu_tests_full.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.log 2>&1
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_ref.log 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.log 2>&1
cat gpurun_out/r2_gpu_tests_full.log; tail -3 gpurun_out/r2_smoke.log
python - <<PY
import json
for f in ['gpurun_out/r2_bench_n1.log','gpurun_out/r2_bench_ref.log','gpurun_out/r2_bench_n2.log']:
    ok=False
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); ok=True
            ex=d.get('extra',{})
            print(f, 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'launch_us', d.get('roofline',{}).get('launch_us'))
            for k in ('train_c4','plan_c3','plan_c5','value_path','steady_state','gcn_layer'):
                if k in ex: print('   ',k, 'value %.4g'%ex[k]['value'], 'ms/step', ex[k].get('ms_per_step'), ex[k].get('grad_allreduce_backend',''), ex[k].get('dp_check',{}).get('ok',''))
    if not ok: print(f,'NO JSON', open(f).read()[-2000:])
PY
