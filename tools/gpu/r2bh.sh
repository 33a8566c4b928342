mkdir -p gpurun_out/r2bh
O=gpurun_out/r2bh
for i in 1 2 3; do
  timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-parity > $O/b_$i.json 2> $O/b_$i.err
  python -c "
import json; d=json.loads(open('$O/b_$i.json').read().strip().split('\n')[-1]); print('run$i', d['ms_per_step'], d['value'], d['clocks'])"
done
