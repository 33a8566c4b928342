mkdir -p gpurun_out/r2z
O=gpurun_out/r2z
timeout 120 python tools/decb_trace.py > $O/decb_trace.log 2>&1; tail -7 $O/decb_trace.log
timeout 120 python tools/dec_trace.py > $O/dec_trace.log 2>&1; tail -4 $O/dec_trace.log
timeout 120 python tools/persist_scaling.py > $O/scaling.log 2>&1; cat $O/scaling.log
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-parity"
timeout 200 $B > $O/c2.json 2> $O/c2.err
MMQG_DEC_BWD_PERSIST=0 timeout 200 $B > $O/c2_off.json 2> $O/c2_off.err
for f in $O/*.json; do echo $f $(python -c "import json;d=json.load(open('$f'));print(d['ms_per_step'], d['gpu_launches']/d['steps'])"); done
