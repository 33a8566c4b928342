mkdir -p gpurun_out/r2al
O=gpurun_out/r2al
timeout 600 python -m pytest tests/test_gpu_convstack.py tests/test_dropin_modules.py -x -q -s > $O/pt.log 2>&1; echo "rc=$?" >> $O/pt.log; grep -E "convstack_|rel err" $O/pt.log | head; tail -15 $O/pt.log
