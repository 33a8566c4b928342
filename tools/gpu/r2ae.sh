mkdir -p gpurun_out/r2ae
O=gpurun_out/r2ae
timeout 600 python -m pytest tests/test_gpu_dp_nccl.py tests/test_gpu_kernels.py -x -q > $O/pt.log 2>&1; echo "rc=$?" >> $O/pt.log; tail -5 $O/pt.log
run() { name=$1; shift; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 4 "$@" > $O/$name.json 2> $O/$name.err; python - <<P
import json
try:
    d = json.loads(open('$O/$name.json').read().strip().split('\n')[-1]); print('$name', d['ms_per_step'], d['value'], d['dp_check'])
except Exception as e:
    print('$name', 'failed', e)
P
}
run n2_fp32 --grad-comm fp32
run n2_bf16 --grad-comm bf16
