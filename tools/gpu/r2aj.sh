mkdir -p gpurun_out/r2aj
O=gpurun_out/r2aj
timeout 400 python -m pytest tests/test_gpu_bf16_mode.py tests/test_gpu_bench_shapes.py -x -q -k "not cfg5" > $O/pt.log 2>&1; echo "rc=$?" >> $O/pt.log; tail -4 $O/pt.log
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-parity"
run() { name=$1; shift; env "$@" timeout 120 $B > $O/$name.json 2> $O/$name.err; echo $name $(python -c "import json;d=json.load(open('$O/$name.json'));print(d['ms_per_step'], d['gpu_launches']/d['steps'])"); }
run perchunk X=1
run old MMQG_HOIST_CHUNKS=0
run perchunk2 X=1
run old2 MMQG_HOIST_CHUNKS=0
timeout 120 python tools/sections.py > $O/sections.log 2>&1; tail -3 $O/sections.log
