mkdir -p gpurun_out/r2ad
O=gpurun_out/r2ad
run() { name=$1; shift; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --steps 20 --warmup 4 > $O/$name.json 2> $O/$name.err; python - <<P
import json
try:
    d = json.loads(open('$O/$name.json').read().strip().split('\n')[-1]); print('$name', d['ms_per_step'], d['value'])
except Exception as e:
    print('$name', 'failed', e)
P
}
run ctas8 NCCL_MAX_CTAS=8
run ctas16 NCCL_MAX_CTAS=16
run ctas4 NCCL_MAX_CTAS=4
run ctas32 NCCL_MAX_CTAS=32
