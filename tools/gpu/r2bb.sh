mkdir -p gpurun_out/r2bb
O=gpurun_out/r2bb
timeout 900 python -m pytest tests/test_gpu_train_step.py "tests/test_gpu_bench_shapes.py::test_cfg2_fp32_tc_step_matches_oracle" -x -q -s > $O/pt.log 2>&1; echo "rc=$?" >> $O/pt.log; grep -E "fp32_tc" $O/pt.log | head -20; tail -15 $O/pt.log
timeout 300 python bench.py --mode fp32_tc --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_tc.json 2> $O/bench_tc.err; cat $O/bench_tc.json; tail -3 $O/bench_tc.err
