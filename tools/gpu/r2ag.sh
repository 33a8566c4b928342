mkdir -p gpurun_out/r2ag
O=gpurun_out/r2ag
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-parity"
run() { name=$1; shift; env "$@" timeout 120 $B > $O/$name.json 2> $O/$name.err; echo $name $(python -c "import json;d=json.load(open('$O/$name.json'));print(d['ms_per_step'], d['gpu_launches']/d['steps'])"); }
run default X=1
run r128_nodefer MMQG_DEC_ROWS=128 MMQG_LH_DEFER=0
run r128_nodefer_pers MMQG_DEC_ROWS=128 MMQG_LH_DEFER=0 MMQG_DEC_BWD_PERSIST=1
run r128_pers MMQG_DEC_ROWS=128 MMQG_DEC_BWD_PERSIST=1
run r128 MMQG_DEC_ROWS=128
run nodefer MMQG_LH_DEFER=0
run nodefer_pers MMQG_LH_DEFER=0 MMQG_DEC_BWD_PERSIST=1
run pers_lh104 MMQG_DEC_BWD_PERSIST=1 MMQG_LH_BYTES=104857600
