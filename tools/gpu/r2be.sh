mkdir -p gpurun_out/r2be
O=gpurun_out/r2be
timeout 900 python -m pytest tests/test_gpu_dp_nccl.py -x -q > $O/pt.log 2>&1; echo "rc=$?" >> $O/pt.log; tail -5 $O/pt.log
