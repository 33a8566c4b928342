mkdir -p gpurun_out/r2bi
O=gpurun_out/r2bi
timeout 500 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline > $O/b4.json 2> $O/b4.err
python -c "
import json; d=json.loads(open('$O/b4.json').read().strip().split('\n')[-1]); print(d['ms_per_step'], d['value'], d['parity'])"
