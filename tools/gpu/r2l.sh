set -x
mkdir -p gpurun_out/r2l
O=gpurun_out/r2l
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-parity"
for cap in 0 48 80 112; do MMQG_LH_CTAS=$cap timeout 120 $B > $O/c2_cap$cap.json 2>/dev/null; done
for by in 16 64; do MMQG_LH_BYTES=$((by<<20)) timeout 120 $B > $O/c2_bytes$by.json 2>/dev/null; done
MMQG_CHUNKS=8 timeout 120 $B > $O/c2_chunks8.json 2>/dev/null
MMQG_CHUNKS=5 timeout 120 $B > $O/c2_chunks5.json 2>/dev/null
MMQG_CHUNKS=4 timeout 120 $B > $O/c2_chunks4.json 2>/dev/null
for f in $O/*.json; do echo $f $(python -c "import json;d=json.load(open('$f'));print(d['ms_per_step'])"); done
