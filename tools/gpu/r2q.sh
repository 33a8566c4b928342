set -x
mkdir -p gpurun_out/r2q
O=gpurun_out/r2q
timeout 400 python -m pytest tests/test_gpu_bf16_mode.py tests/test_gpu_bench_shapes.py tests/test_gpu_lstm_seq.py -x -q -s -k "not cfg5 and not cfg4" > $O/pt.log 2>&1; echo "rc=$?" >> $O/pt.log; tail -5 $O/pt.log
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-parity"
timeout 120 $B > $O/c2.json 2> $O/c2.err
MMQG_FWD2=0 timeout 120 $B > $O/c2_fwd1.json 2> $O/c2_fwd1.err
MMQG_CHUNKS=8 timeout 120 $B > $O/c2_ch8.json 2> $O/c2_ch8.err
timeout 120 python tools/ktrace.py --graph > $O/ktrace.log 2>&1
timeout 200 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu-baseline > $O/c5.json 2> $O/c5.err
for f in $O/*.json; do echo $f $(python -c "import json;d=json.load(open('$f'));print(d['ms_per_step'], d['gpu_launches']/d['steps'], d.get('parity'))"); done
cat $O/ktrace.log | tail -45
