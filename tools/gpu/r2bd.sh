mkdir -p gpurun_out/r2bd
O=gpurun_out/r2bd
CMD="python tools/bench_convstack.py 2560"
ncu --set full --import-source on --clock-control none -k regex:"conv3_" -s 12 -c 12 -o $O/ncu_conv3 -f $CMD > $O/ncu.log 2>&1
tail -2 $O/ncu.log; ls -la $O
