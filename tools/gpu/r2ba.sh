mkdir -p gpurun_out/r2ba
O=gpurun_out/r2ba
NOACQ=$PWD/multi-modal-qg_b200/mmqg/libmmqg_noacq.so
for v in base noacq base2 noacq2; do
  if [[ $v == noacq* ]]; then export MMQG_LIB=$NOACQ; else unset MMQG_LIB; fi
  timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/b_$v.json 2> $O/b_$v.err
  python -c "
import json; d=json.loads(open('$O/b_$v.json').read().strip().split('\n')[-1]); print('$v', d['ms_per_step'], d['value'], d['parity']['loss_rel'], d['parity']['worst_grad_rel'], d['roofline'].get('avg_launch_us'))"
done
export MMQG_LIB=$NOACQ
timeout 900 python -m pytest tests/test_gpu_bf16_mode.py tests/test_gpu_bench_shapes.py tests/test_gpu_dec_bwd_persist.py -k "not cfg4" -x -q > $O/pt.log 2>&1; echo "rc=$?" >> $O/pt.log; tail -3 $O/pt.log
