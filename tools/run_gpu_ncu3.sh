mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $BENCH > gpurun_out/plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 54 -c 18 -o gpurun_out/prof_conv -f $BENCH > gpurun_out/ncu_conv.log 2>&1
echo "ncu conv exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
