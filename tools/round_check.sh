#!/bin/bash
# Everything the round-end driver runs, plus the ncu captures the profiles/ summaries come from (2 B200s, ~6 minutes).
# Usage (from the repo root, through gpurun --gpus 2):  bash tools/round_check.sh
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -6 > gpurun_out/gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.log 2>&1
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.log 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_plan_c3.csv python bench.py --workload plan --plan-eager --no-cpu-baseline --steps 3 --warmup 1 > gpurun_out/plan_ncu.log 2>&1
timeout 250 ncu --set full --clock-control none --import-source on -k regex:graph_forward_tp -s 3 -c 1 -o gpurun_out/tp_b1m -f python tools/prof_graph.py 1048576 > gpurun_out/ncu_tp.log 2>&1
timeout 200 python tools/quick_time.py > gpurun_out/qt.log 2>&1
timeout 200 python tools/bwd_time.py > gpurun_out/bwd_time.log 2>&1
cat gpurun_out/gpu_tests.log gpurun_out/smoke.log gpurun_out/qt.log gpurun_out/bwd_time.log
