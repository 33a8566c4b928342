set -x
mkdir -p gpurun_out/r2j
O=gpurun_out/r2j
timeout 400 python -m pytest tests/test_gpu_bf16_mode.py tests/test_gpu_bench_shapes.py tests/test_gpu_train_step.py tests/test_gpu_adam.py -x -q -s -k "not cfg5" > $O/pt.log 2>&1; echo "rc=$?" >> $O/pt.log; tail -5 $O/pt.log
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-parity"
timeout 120 $B > $O/c2.json 2> $O/c2.err
MMQG_LH_DEFER=0 timeout 120 $B > $O/c2_nodefer.json 2> $O/c2_nodefer.err
MMQG_DEC_PERSIST=0 timeout 120 $B > $O/c2_dec_off.json 2> $O/c2_dec_off.err
timeout 120 python tools/sections.py > $O/sections.log 2>&1
timeout 120 python tools/dec_trace.py > $O/dec_trace.log 2>&1
timeout 200 python bench.py --config 4 --steps 8 --warmup 3 --no-cpu-baseline --no-parity > $O/c4.json 2> $O/c4.err
for f in $O/*.json; do echo $f $(python -c "import json;d=json.load(open('$f'));print(d['ms_per_step'], d['gpu_launches']/d['steps'])"); done
cat $O/sections.log; tail -5 $O/dec_trace.log
