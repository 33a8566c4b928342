mkdir -p gpurun_out/r2z
O=gpurun_out/r2z
timeout 120 python tools/sections.py > $O/sections.log 2>&1; cat $O/sections.log
MMQG_DEC_BWD_PERSIST=0 timeout 120 python tools/sections.py > $O/sections_off.log 2>&1; cat $O/sections_off.log
