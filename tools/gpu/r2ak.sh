mkdir -p gpurun_out/r2ak
O=gpurun_out/r2ak
timeout 600 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_bf16_mode.py tests/test_gpu_adam.py -x -q -s > $O/pt.log 2>&1; echo "rc=$?" >> $O/pt.log; grep "fp32 + dropout" $O/pt.log; tail -4 $O/pt.log
