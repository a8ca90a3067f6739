#!/bin/bash
# Everything the round-end driver runs, plus the ncu captures the profiles/ summaries come from (one B200, ~3 minutes).
# Usage (from the repo root, through gpurun):  bash tools/round_check.sh
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x --timeout 150 2>&1 | tail -3 > gpurun_out/gpu_tests.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 300 --csv --log-file gpurun_out/launches_b4096.csv python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:graph_forward_tc -s 5 -c 1 -o gpurun_out/tc5_b4096 -f python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:graph_forward_tc -s 3 -c 1 -o gpurun_out/tc5_b1m -f python tools/prof_graph.py 1048576 > gpurun_out/ncu3.log 2>&1
timeout 250 python bench.py > gpurun_out/bench_full.log 2>&1
timeout 150 python bench.py --impl reference > gpurun_out/bench_ref.log 2>&1
timeout 100 python bench.py --workload plan --no-cpu-baseline --steps 50 > gpurun_out/bench_plan.log 2>&1
timeout 100 python bench.py --workload plan --no-cpu-baseline --steps 20 --humans 20 --depth 3 --roots 2048 --speed-samples 5 --rotation-samples 16 > gpurun_out/bench_plan_c5.log 2>&1
timeout 100 python tools/quick_time.py > gpurun_out/tc_qt_full.log 2>&1
cat gpurun_out/gpu_tests.log gpurun_out/smoke.log gpurun_out/tc_qt_full.log
