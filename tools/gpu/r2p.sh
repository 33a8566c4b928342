set -x
mkdir -p gpurun_out/r2p
O=gpurun_out/r2p
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-graph"
$CMD > $O/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 420 --csv --log-file $O/launches.csv $CMD > $O/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lstm_seq_bwd4_kernel -s 20 -c 3 -o $O/prof_bwd4 $CMD > $O/ncu_bwd4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lstm_seq_fwd_kernel -s 20 -c 3 -o $O/prof_fwd $CMD > $O/ncu_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dec_seq_fwd_kernel -s 2 -c 1 -o $O/prof_dec $CMD > $O/ncu_dec.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_persist_kernel -s 60 -c 12 -o $O/prof_persist $CMD > $O/ncu_persist.log 2>&1
ncu --set full --clock-control none -k regex:attn_ -s 40 -c 4 -o $O/prof_attn $CMD > $O/ncu_attn.log 2>&1
ls -la $O
