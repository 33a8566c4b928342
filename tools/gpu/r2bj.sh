mkdir -p gpurun_out/r2bj
O=gpurun_out/r2bj
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 8 --steps 20 --warmup 5 > $O/n8.json 2> $O/n8.err
python -c "
import json; d=json.loads(open('$O/n8.json').read().strip().split('\n')[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['dp_check']['worst_grad_rel'], d['config'].get('grad_reduce'), d['clocks'])"
tail -3 $O/n8.err
