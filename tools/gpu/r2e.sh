set -x
mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
timeout 300 python -m pytest tests/test_gpu_vocab_nll.py tests/test_gpu_gemm_tc.py -x -q -s > $O/pytest_vocab.log 2>&1; echo "rc=$?" >> $O/pytest_vocab.log
tail -8 $O/pytest_vocab.log
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-parity"
for bytes in 16 32 64; do for steps in 0 2 3; do for tail in 1; do
  MMQG_LH_BYTES=$((bytes<<20)) MMQG_LH_STEPS=$steps MMQG_LH_TAIL=$tail timeout 120 $B > $O/c2_b${bytes}_s${steps}_t${tail}.json 2>/dev/null
done; done; done
MMQG_LH_TAIL=0 timeout 120 $B > $O/c2_default_t0.json 2>/dev/null
MMQG_LH_TAIL=2 timeout 120 $B > $O/c2_default_t2.json 2>/dev/null
B4="python bench.py --config 4 --steps 8 --warmup 3 --no-cpu-baseline --no-parity"
timeout 200 $B4 > $O/c4_default.json 2>/dev/null
MMQG_LH_MINROWS=1024 timeout 200 $B4 > $O/c4_min1024.json 2>/dev/null
MMQG_LH_MINROWS=256 timeout 200 $B4 > $O/c4_min256.json 2>/dev/null
for f in $O/*.json; do echo $f $(python -c "import json;d=json.load(open('$f'));print(d['ms_per_step'])"); done
