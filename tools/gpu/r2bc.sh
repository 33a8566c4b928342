mkdir -p gpurun_out/r2bc
O=gpurun_out/r2bc
timeout 900 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_bf16_mode.py tests/test_gpu_adam.py tests/test_gpu_decoder_nonattn.py tests/test_dropin_modules.py -x -q -s > $O/pt.log 2>&1; echo "rc=$?" >> $O/pt.log; grep -E "varlen" $O/pt.log | head; tail -12 $O/pt.log
