mkdir -p gpurun_out/r2at
O=gpurun_out/r2at
timeout 600 python -m pytest tests/test_gpu_convstack.py -x -q -s > $O/pt.log 2>&1; echo "rc=$?" >> $O/pt.log; grep -E "rel err" $O/pt.log | head -3; tail -12 $O/pt.log
timeout 200 python tools/bench_convstack.py 2560 > $O/conv_2560.json 2> $O/conv.err; cat $O/conv_2560.json; tail -3 $O/conv.err
CMD="python tools/bench_convstack.py 2560"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"conv|bn_|maxpool" -s 60 -c 24 --csv --log-file $O/conv_launches.csv $CMD > $O/ncu_conv.log 2>&1
tail -1 $O/ncu_conv.log
