mkdir -p gpurun_out/r2y
O=gpurun_out/r2y
timeout 150 python -m pytest tests/test_gpu_bf16_mode.py -x -q -s -k "test_bf16_step_matches_rounded_oracle" > $O/pt_small.log 2>&1; echo "rc=$?" >> $O/pt_small.log; tail -12 $O/pt_small.log
timeout 400 python -m pytest tests/test_gpu_bf16_mode.py tests/test_gpu_bench_shapes.py -x -q -s -k "not cfg5 and not cfg4" > $O/pt_mid.log 2>&1; echo "rc=$?" >> $O/pt_mid.log; tail -12 $O/pt_mid.log
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
timeout 200 $B > $O/c2.json 2> $O/c2.err
MMQG_DEC_BWD_PERSIST=0 timeout 200 $B --no-parity > $O/c2_off.json 2> $O/c2_off.err
for f in $O/*.json; do echo $f $(python -c "import json;d=json.load(open('$f'));print(d['ms_per_step'], d['gpu_launches']/d['steps'], d.get('parity'))"); done
timeout 120 python tools/sections.py > $O/sections.log 2>&1; cat $O/sections.log
