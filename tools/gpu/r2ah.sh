mkdir -p gpurun_out/r2ah
O=gpurun_out/r2ah
timeout 400 python -m pytest tests/test_gpu_bf16_mode.py tests/test_gpu_bench_shapes.py tests/test_gpu_lstm_seq.py tests/test_gpu_decoder_nonattn.py -x -q -k "not cfg5 and not cfg4" > $O/pt.log 2>&1; echo "rc=$?" >> $O/pt.log; tail -4 $O/pt.log
timeout 120 python tools/persist_scaling.py > $O/scaling.log 2>&1; cat $O/scaling.log
timeout 120 python tools/trace_lstm_bwd.py > $O/trace_bwd.log 2>&1; head -16 $O/trace_bwd.log
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-parity"
timeout 200 $B > $O/c2.json 2> $O/c2.err
for f in $O/*.json; do echo $f $(python -c "import json;d=json.load(open('$f'));print(d['ms_per_step'], d['gpu_launches']/d['steps'])"); done
