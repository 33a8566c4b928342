mkdir -p gpurun_out/r2bf
O=gpurun_out/r2bf
export MMQG_FWD_MC=1
timeout 300 python -m pytest tests/test_gpu_lstm_seq.py tests/test_gpu_bf16_mode.py -x -q > $O/pt.log 2>&1; echo "rc=$?" >> $O/pt.log; tail -3 $O/pt.log
for v in mc base mc2 base2; do
  if [[ $v == mc* ]]; then export MMQG_FWD_MC=1; else export MMQG_FWD_MC=0; fi
  timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/b_$v.json 2> $O/b_$v.err
  python -c "
import json; d=json.loads(open('$O/b_$v.json').read().strip().split('\n')[-1]); print('$v', d['ms_per_step'], d['value'], d['parity']['loss_rel'], d['parity']['worst_grad_rel'], d['roofline'].get('avg_launch_us'))"
done
