set -x
mkdir -p gpurun_out/r2g
O=gpurun_out/r2g
# persistent decoder forward: small shapes first (each under its own timeout: a hang must not eat the call)
timeout 120 python -m pytest tests/test_gpu_bf16_mode.py -x -q -s -k "test_bf16_step_matches_rounded_oracle" > $O/pt_small.log 2>&1; echo "rc=$?" >> $O/pt_small.log; tail -6 $O/pt_small.log
timeout 300 python -m pytest tests/test_gpu_bf16_mode.py tests/test_gpu_bench_shapes.py -x -q -s -k "not cfg4 and not cfg5" > $O/pt_mid.log 2>&1; echo "rc=$?" >> $O/pt_mid.log; tail -12 $O/pt_mid.log
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-parity"
timeout 120 $B > $O/c2_dec64.json 2> $O/c2_dec64.err
MMQG_DEC_ROWS=128 timeout 120 $B > $O/c2_dec128.json 2> $O/c2_dec128.err
MMQG_DEC_PERSIST=0 timeout 120 $B > $O/c2_dec_off.json 2> $O/c2_dec_off.err
timeout 120 python tools/sections.py > $O/sections.log 2>&1
MMQG_DEC_ROWS=128 timeout 120 python tools/sections.py > $O/sections128.log 2>&1
for f in $O/*.json; do echo $f $(python -c "import json;d=json.load(open('$f'));print(d['ms_per_step'])"); done
cat $O/sections.log $O/sections128.log
