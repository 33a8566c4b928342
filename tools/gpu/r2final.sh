mkdir -p gpurun_out/r2final2
O=gpurun_out/r2final2
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pt.log 2>&1; echo "rc=$?" >> $O/pt.log; tail -4 $O/pt.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log; tail -6 $O/smoke.log
timeout 400 python bench.py > $O/bench.json 2> $O/bench.err; python -c "
import json; d=json.loads(open('$O/bench.json').read().strip().split('\n')[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['parity']['ok'], d['roofline']['frac'], d['cpu_baseline']['value'], d['gpu_launches'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 600 $O/bench_ref.json
