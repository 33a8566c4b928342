set -x
mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
timeout 300 python -m pytest tests/test_gpu_lstm_seq.py tests/test_gpu_decoder_nonattn.py tests/test_dropin_modules.py -x -q -s > $O/pytest_new.log 2>&1; echo "rc=$?" >> $O/pytest_new.log
tail -25 $O/pytest_new.log
timeout 900 python -m pytest tests -m gpu -q -s > $O/pytest_all.log 2>&1; echo "rc=$?" >> $O/pytest_all.log
tail -6 $O/pytest_all.log
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err
for f in $O/*.json; do echo $f $(python -c "import json;d=json.load(open('$f'));print(d['ms_per_step'], d.get('parity'))"); done
