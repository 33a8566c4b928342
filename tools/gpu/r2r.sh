set -x
mkdir -p gpurun_out/r2r
O=gpurun_out/r2r
timeout 120 python tools/trace_lstm2.py > $O/trace2.log 2>&1
cat $O/trace2.log
timeout 120 python tools/trace_lstm.py > $O/trace1.log 2>&1
MMQG_FWD2=0 timeout 120 python tools/trace_lstm.py > $O/trace1_old.log 2>&1
cat $O/trace1_old.log
