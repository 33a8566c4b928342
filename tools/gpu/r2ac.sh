mkdir -p gpurun_out/r2ac
O=gpurun_out/r2ac
nvidia-smi -L | wc -l
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --steps 30 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err; echo "rc=$?"
tail -3 $O/bench_n8.err
python -c "
import json;d=json.load(open('$O/bench_n8.json'));print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('dp_check'))"
timeout 300 python -m pytest tests/test_gpu_dp_nccl.py -x -q > $O/pt_dp.log 2>&1; tail -3 $O/pt_dp.log
