mkdir -p gpurun_out/r2u
O=gpurun_out/r2u
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-parity"
timeout 120 $B > $O/c2_fwd2.json 2> $O/c2_fwd2.err
MMQG_NOCOOP=1 timeout 120 $B > $O/c2_fwd2_nocoop.json 2> $O/c2_fwd2_nocoop.err
MMQG_NOCOOP=1 MMQG_CHUNKS=8 timeout 120 $B > $O/c2_fwd2_nocoop_ch8.json 2> $O/c2_fwd2_nocoop_ch8.err
MMQG_FWD2=0 timeout 120 $B > $O/c2_fwd1.json 2> $O/c2_fwd1.err
for f in $O/*.json; do echo $f $(python -c "import json;d=json.load(open('$f'));print(d['ms_per_step'])"); done
MMQG_NOCOOP=1 timeout 120 python tools/ktrace.py --graph > $O/ktrace_nocoop.log 2>&1
head -22 $O/ktrace_nocoop.log; tail -2 $O/ktrace_nocoop.log
