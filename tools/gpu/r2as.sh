mkdir -p gpurun_out/r2as
O=gpurun_out/r2as
CMD="python tools/bench_convstack.py 2560"
ncu --set full --import-source on --clock-control none -k regex:"conv3_bwd_w" -s 4 -c 4 -o $O/ncu_bwd_w -f $CMD > $O/ncu.log 2>&1
tail -2 $O/ncu.log; ls -la $O
