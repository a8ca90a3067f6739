mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gcn_policy.py tests/test_planner.py -m gpu -q -x --timeout 200 2>&1 | tail -4 > gpurun_out/r2_gcn_tests.log
timeout 250 ncu --set full --clock-control none --import-source on -k regex:graph_forward_tp -s 5 -c 1 -o gpurun_out/r2_tp_b4096 -f python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_tp4096.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 160 --csv --log-file gpurun_out/r2_launches_train.csv python bench.py --workload train --batch 8192 --humans 10 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_train_ncu.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 300 --csv --log-file gpurun_out/r2_launches_b4096.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_b_ncu.log 2>&1
cat gpurun_out/r2_gcn_tests.log; tail -2 gpurun_out/r2_ncu_tp4096.log
