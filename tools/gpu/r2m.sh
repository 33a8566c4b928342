set -x
mkdir -p gpurun_out/r2m
O=gpurun_out/r2m
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_dp_nccl.py -x -q -s > $O/pt_dp.log 2>&1; echo "rc=$?" >> $O/pt_dp.log; tail -15 $O/pt_dp.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 30 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "rc=$?"
tail -3 $O/bench_n2.err
python -c "
import json;d=json.load(open('$O/bench_n2.json'));print(d['value'], d['ms_per_step'], d.get('dp_check'))"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/ref_n2.json 2> $O/ref_n2.err; cat $O/ref_n2.json | cut -c1-400
