mkdir -p gpurun_out/r2v
O=gpurun_out/r2v
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-parity"
run() { name=$1; shift; env "$@" timeout 120 $B > $O/$name.json 2> $O/$name.err; echo $name $(python -c "import json;d=json.load(open('$O/$name.json'));print(d['ms_per_step'])"); }
run f2_c6 MMQG_FWD2=1
run f2_c8_g48 MMQG_CHUNKS=8 MMQG_GCAP=48
run f2_c12_g48 MMQG_CHUNKS=12 MMQG_GCAP=48
run f2_c12_g32 MMQG_CHUNKS=12 MMQG_GCAP=32
run f2_c12_g0 MMQG_CHUNKS=12
run f2_c6_g48 MMQG_GCAP=48
run f1_c6 MMQG_FWD2=0
run f1_c12 MMQG_FWD2=0 MMQG_CHUNKS=12
MMQG_CHUNKS=12 MMQG_GCAP=48 timeout 120 python tools/ktrace.py --graph > $O/ktrace_c12_g48.log 2>&1
head -40 $O/ktrace_c12_g48.log; tail -2 $O/ktrace_c12_g48.log
