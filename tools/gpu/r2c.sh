set -x
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
timeout 400 python -m pytest tests/test_gpu_bf16_mode.py tests/test_gpu_bench_shapes.py -x -q -s -k "not cfg4 and not cfg5" > $O/pytest_bwd4.log 2>&1; echo "rc=$?" >> $O/pytest_bwd4.log
tail -4 $O/pytest_bwd4.log
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-parity > $O/bench_bwd4.json 2> $O/bench_bwd4.err
timeout 120 python tools/ktrace.py --graph > $O/ktrace.log 2>&1
grep -h '"value"' $O/*.json | cut -c1-120
tail -3 $O/ktrace.log
