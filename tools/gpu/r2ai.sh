mkdir -p gpurun_out/r2ai
O=gpurun_out/r2ai
timeout 120 python tools/persist_scaling.py > $O/scaling.log 2>&1; cat $O/scaling.log
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-parity"
timeout 200 $B > $O/c2.json 2> $O/c2.err
for f in $O/*.json; do echo $f $(python -c "import json;d=json.load(open('$f'));print(d['ms_per_step'], d['gpu_launches']/d['steps'])"); done
