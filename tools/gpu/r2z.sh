mkdir -p gpurun_out/r2z
O=gpurun_out/r2z
timeout 300 python -m pytest tests/test_gpu_bf16_mode.py tests/test_gpu_bench_shapes.py -x -q -k "not cfg5 and not cfg4" > $O/pt_mid.log 2>&1; echo "rc=$?" >> $O/pt_mid.log; tail -4 $O/pt_mid.log
timeout 120 python tools/decb_trace.py > $O/decb_trace.log 2>&1; tail -7 $O/decb_trace.log
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-parity"
timeout 200 $B > $O/c2.json 2> $O/c2.err
for f in $O/*.json; do echo $f $(python -c "import json;d=json.load(open('$f'));print(d['ms_per_step'], d['gpu_launches']/d['steps'])"); done
