mkdir -p gpurun_out/r2bg
O=gpurun_out/r2bg
timeout 120 python tools/sections.py > $O/sections_base.log 2>&1; head -3 $O/sections_base.log
MMQG_DEBUG_NOSAVE=1 timeout 120 python tools/sections.py > $O/sections_nosave.log 2>&1; head -3 $O/sections_nosave.log
