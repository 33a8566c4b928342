mkdir -p gpurun_out/r2s
O=gpurun_out/r2s
for dbg in 0 1 2 3 4 5; do
  echo "=== MMQG_DBG_FWD=$dbg"
  MMQG_DBG_FWD=$dbg timeout 120 python tools/trace_lstm2.py > $O/trace_dbg$dbg.log 2>&1
  head -15 $O/trace_dbg$dbg.log
done
echo "=== p=0"
DROP_P=0 timeout 120 python tools/trace_lstm2.py > $O/trace_p0.log 2>&1
head -15 $O/trace_p0.log
