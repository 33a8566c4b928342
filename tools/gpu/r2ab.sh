mkdir -p gpurun_out/r2ab
O=gpurun_out/r2ab
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-parity"
run() { name=$1; shift; env "$@" timeout 120 $B > $O/$name.json 2> $O/$name.err; echo $name $(python -c "import json;d=json.load(open('$O/$name.json'));print(d['ms_per_step'])"); }
run default X=1
run pers MMQG_DEC_BWD_PERSIST=1
run pers_lh48 MMQG_DEC_BWD_PERSIST=1 MMQG_LH_BYTES=50331648
run pers_lh64 MMQG_DEC_BWD_PERSIST=1 MMQG_LH_BYTES=67108864
run pers_lh104 MMQG_DEC_BWD_PERSIST=1 MMQG_LH_BYTES=104857600
run pers_lh16 MMQG_DEC_BWD_PERSIST=1 MMQG_LH_BYTES=16777216
MMQG_DEC_BWD_PERSIST=1 MMQG_LH_BYTES=104857600 timeout 120 python tools/sections.py > $O/sections_lh104.log 2>&1; grep "BPTT loop" $O/sections_lh104.log
MMQG_DEC_BWD_PERSIST=1 MMQG_LH_BYTES=50331648 timeout 120 python tools/sections.py > $O/sections_lh48.log 2>&1; grep "BPTT loop" $O/sections_lh48.log
