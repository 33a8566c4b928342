set -x
mkdir -p gpurun_out/r2n
O=gpurun_out/r2n
timeout 300 python -X faulthandler -u -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.log 2>&1; echo "rc=$?"
tail -30 $O/bench_n2.log
