set -x
mkdir -p gpurun_out/r2a
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a/pytest.log
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/r2a/bench_c2.json 2> gpurun_out/r2a/bench_c2.err
timeout 400 python bench.py --config 4 --steps 10 --warmup 3 --cpu-samples 4 > gpurun_out/r2a/bench_c4.json 2> gpurun_out/r2a/bench_c4.err
timeout 300 python bench.py --config 5 --steps 10 --warmup 3 > gpurun_out/r2a/bench_c5.json 2> gpurun_out/r2a/bench_c5.err
timeout 120 python tools/sections.py > gpurun_out/r2a/sections.log 2>&1
timeout 120 python tools/ktrace.py --graph > gpurun_out/r2a/ktrace.log 2>&1
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-graph > gpurun_out/r2a/plain_attn.log 2>&1 && \
timeout 600 ncu --set full --cache-control none --clock-control none -k regex:attn_ -s 40 -c 6 -o gpurun_out/r2a/prof_attn_nocc python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-graph > gpurun_out/r2a/ncu_attn.log 2>&1
ls -la gpurun_out/r2a
