set -x
mkdir -p gpurun_out/r2d
O=gpurun_out/r2d
timeout 300 python -m pytest tests/test_gpu_vocab_nll.py -x -q -s > $O/pytest_vocab.log 2>&1; echo "rc=$?" >> $O/pytest_vocab.log
tail -15 $O/pytest_vocab.log
timeout 900 python -m pytest tests -m gpu -x -q -s --deselect tests/test_gpu_vocab_nll.py > $O/pytest_all.log 2>&1; echo "rc=$?" >> $O/pytest_all.log
tail -5 $O/pytest_all.log
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err
timeout 300 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err
timeout 200 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c5.json 2> $O/bench_c5.err
timeout 120 python tools/sections.py > $O/sections.log 2>&1
grep -h -o '"value": [0-9.]*, "unit": "samples/s", "n_gpus"' $O/*.json
cat $O/sections.log
