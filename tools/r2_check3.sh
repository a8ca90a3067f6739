#!/bin/bash
# round-2 check on 2 GPUs: full GPU test suite, planner launch list, value-head ncu captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -25 > gpurun_out/r2_gpu_tests.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_plan_c3.csv python bench.py --workload plan --plan-eager --no-cpu-baseline --steps 3 --warmup 1 > gpurun_out/r2_plan_ncu.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:value_head_tc -s 3 -c 1 -o gpurun_out/r2_vh_b4096 -f python tools/prof_value.py 4096 > gpurun_out/r2_ncu_vh1.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:value_head_tc -s 3 -c 1 -o gpurun_out/r2_vh_b1m -f python tools/prof_value.py 1048576 > gpurun_out/r2_ncu_vh2.log 2>&1
timeout 200 python tools/quick_time.py > gpurun_out/r2_qt.log 2>&1
cat gpurun_out/r2_gpu_tests.log; cat gpurun_out/r2_qt.log
