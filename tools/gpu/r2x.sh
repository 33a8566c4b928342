mkdir -p gpurun_out/r2x
O=gpurun_out/r2x
timeout 120 python tools/trace_lstm_bwd.py > $O/trace_bwd.log 2>&1; cat $O/trace_bwd.log
