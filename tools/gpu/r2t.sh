mkdir -p gpurun_out/r2t
O=gpurun_out/r2t
for dbg in 0 8 10; do
  echo "=== fwd2 MMQG_DBG_FWD=$dbg"
  MMQG_DBG_FWD=$dbg timeout 120 python tools/trace_lstm2.py > $O/trace_dbg$dbg.log 2>&1
  head -17 $O/trace_dbg$dbg.log
done
for dbg in 0 8 10 12; do
  echo "=== old kernel MMQG_DBG_FWD=$dbg"
  MMQG_FWD2=0 MMQG_DBG_FWD=$dbg timeout 120 python tools/persist_scaling.py > $O/scal_dbg$dbg.log 2>&1
  cat $O/scal_dbg$dbg.log
done
