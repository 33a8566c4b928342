mkdir -p gpurun_out/r2av
O=gpurun_out/r2av
CMD="python bench.py --config 5 --steps 2 --warmup 3 --no-parity --no-cpu-baseline --no-graph"
timeout 300 $CMD > $O/b5.json 2> $O/b5.err; python -c "
import json; d=json.loads(open('$O/b5.json').read().strip().split('\n')[-1]); print(d['ms_per_step'], d['value'], d['gpu_launches'])"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 1430 -c 450 --csv --log-file $O/launches5.csv $CMD > $O/ncu5.log 2>&1
tail -1 $O/ncu5.log; wc -l $O/launches5.csv
