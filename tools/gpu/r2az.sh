mkdir -p gpurun_out/r2az
O=gpurun_out/r2az
N=$1
run() { name=$1; shift; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 20 --warmup 4 "$@" > $O/$name.json 2> $O/$name.err; python - <<P
import json
try:
    d = json.loads(open('$O/$name.json').read().strip().split('\n')[-1]); print('$name', d['ms_per_step'], d['value'], d['dp_check']['worst_grad_rel'], d['config'].get('grad_reduce'))
except Exception as e:
    print('$name', 'failed', e); print(open('$O/$name.err').read()[-1500:])
P
}
run n${N}_default
MMQG_MM_CTAS=4 run n${N}_mm4
