mkdir -p gpurun_out/r2ax
O=gpurun_out/r2ax
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 150 $TR --master-port 29601 tools/check_dp.py --mode fp32 --reduce multimem > $O/chk_fp32.log 2>&1; echo "rc=$?" >> $O/chk_fp32.log; grep -E "dp world|rc=|Error|error" $O/chk_fp32.log | head -8
timeout 150 $TR --master-port 29602 tools/check_dp.py --mode bf16 --reduce multimem > $O/chk_bf16.log 2>&1; echo "rc=$?" >> $O/chk_bf16.log; grep -E "dp world|rc=|Error|error" $O/chk_bf16.log | head -8
run() { name=$1; shift; timeout 200 $TR --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 4 "$@" > $O/$name.json 2> $O/$name.err; python - <<P
import json
try:
    d = json.loads(open('$O/$name.json').read().strip().split('\n')[-1]); print('$name', d['ms_per_step'], d['value'], d['dp_check']['worst_grad_rel'])
except Exception as e:
    print('$name', 'failed', e); print(open('$O/$name.err').read()[-1500:])
P
}
run n2_mm --reduce multimem
run n2_nccl --reduce overlap
