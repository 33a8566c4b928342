set -x
mkdir -p gpurun_out/r2h
O=gpurun_out/r2h
timeout 120 python tools/dec_trace.py > $O/dec_trace.log 2>&1
MMQG_LOSS_OVERLAP=0 timeout 120 python tools/dec_trace.py > $O/dec_trace_nolh.log 2>&1
MMQG_LOSS_OVERLAP=0 timeout 120 python tools/sections.py > $O/sections_nolh.log 2>&1
MMQG_LOSS_OVERLAP=0 MMQG_DEC_ROWS=128 timeout 120 python tools/dec_trace.py > $O/dec_trace_nolh128.log 2>&1
cat $O/dec_trace_nolh.log; cat $O/sections_nolh.log; tail -3 $O/dec_trace.log; tail -8 $O/dec_trace_nolh128.log
