mkdir -p gpurun_out/r2aa
O=gpurun_out/r2aa
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_all.log 2>&1; echo "rc=$?" >> $O/pytest_all.log; tail -6 $O/pytest_all.log
B="--steps 30 --warmup 5"
timeout 300 python bench.py $B > $O/c2.json 2> $O/c2.err
for f in $O/*.json; do echo $f $(python -c "import json;d=json.load(open('$f'));print(d['ms_per_step'], d['value'], d['gpu_launches']/d['steps'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline'])"); done
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-graph"
MMQG_NCU=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 1750 -c 420 --csv --log-file $O/launches.csv $CMD > $O/ncu_list.log 2>&1
tail -2 $O/ncu_list.log
MMQG_NCU=1 ncu --set full --clock-control none --import-source on -k regex:lstm_seq_bwd4_kernel -s 20 -c 3 -o $O/prof_bwd4 $CMD > $O/ncu_bwd4.log 2>&1
tail -2 $O/ncu_bwd4.log
MMQG_NCU=1 MMQG_DEC_BWD_PERSIST=1 ncu --set full --clock-control none --import-source on -k regex:dec_seq_bwd_kernel -s 2 -c 1 -o $O/prof_decb $CMD > $O/ncu_decb.log 2>&1
tail -2 $O/ncu_decb.log
ls -la $O
