set -x
mkdir -p gpurun_out/r2i
O=gpurun_out/r2i
timeout 300 python -m pytest tests/test_gpu_bf16_mode.py tests/test_gpu_bench_shapes.py -x -q -s -k "not cfg4 and not cfg5" > $O/pt_mid.log 2>&1; echo "rc=$?" >> $O/pt_mid.log; tail -4 $O/pt_mid.log
MMQG_LOSS_OVERLAP=0 timeout 120 python tools/dec_trace.py > $O/dec_trace_nolh.log 2>&1
tail -8 $O/dec_trace_nolh.log
