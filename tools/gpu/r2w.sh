mkdir -p gpurun_out/r2w
O=gpurun_out/r2w
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_all.log 2>&1; echo "rc=$?" >> $O/pytest_all.log; tail -6 $O/pytest_all.log
B="--steps 30 --warmup 5 --no-cpu-baseline"
timeout 200 python bench.py $B > $O/c2.json 2> $O/c2.err
timeout 200 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu-baseline > $O/c5.json 2> $O/c5.err
MMQG_DEC_PERSIST=0 timeout 200 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu-baseline --no-parity > $O/c5_nodp.json 2> $O/c5_nodp.err
timeout 300 python bench.py --config 4 --steps 8 --warmup 3 --no-cpu-baseline > $O/c4.json 2> $O/c4.err
for f in $O/*.json; do echo $f $(python -c "import json;d=json.load(open('$f'));print(d['ms_per_step'], d['value'], d['gpu_launches']/d['steps'], d.get('parity'))"); done
tail -3 $O/c5.err
