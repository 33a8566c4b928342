mkdir -p gpurun_out/r2aw
O=gpurun_out/r2aw
timeout 900 python -m pytest tests/test_gpu_bf16_mode.py tests/test_gpu_sampling.py tests/test_gpu_kernels.py tests/test_gpu_bench_shapes.py -k "not cfg4" -x -q -s > $O/pt.log 2>&1; echo "rc=$?" >> $O/pt.log; grep -E "cfg-5|tokens equal" $O/pt.log | head; tail -5 $O/pt.log
timeout 300 python bench.py --config 5 --steps 10 --warmup 3 > $O/b5.json 2> $O/b5.err; python -c "
import json; d=json.loads(open('$O/b5.json').read().strip().split('\n')[-1]); print(d['ms_per_step'], d['value'], d['gpu_launches'], d.get('parity'))"
