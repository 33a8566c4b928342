mkdir -p gpurun_out/r2ay
O=gpurun_out/r2ay
run() { name=$1; shift; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --steps 20 --warmup 4 "$@" > $O/$name.json 2> $O/$name.err; python - <<P
import json
try:
    d = json.loads(open('$O/$name.json').read().strip().split('\n')[-1]); print('$name', d['ms_per_step'], d['value'], d['dp_check']['worst_grad_rel'])
except Exception as e:
    print('$name', 'failed', e); print(open('$O/$name.err').read()[-1500:])
P
}
MMQG_MM_CTAS=16 run n8_mm16 --reduce multimem
MMQG_MM_CTAS=8 run n8_mm8 --reduce multimem
MMQG_MM_CTAS=32 run n8_mm32 --reduce multimem
