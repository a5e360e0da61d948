mkdir -p gpurun_out
for t in test_gpu_conv test_gpu_net; do
  timeout 300 python -m pytest tests/$t.py -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/$t.log 2>&1
  echo "$t exit $?" >> gpurun_out/summary.txt
done
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 300 $BENCH > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launch.log 2>&1
echo "launchlist exit $?" >> gpurun_out/summary.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 54 -c 18 -o gpurun_out/prof_conv -f $BENCH > gpurun_out/ncu_conv.log 2>&1
echo "ncu conv exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt gpurun_out/bench.json; tail -3 gpurun_out/test_gpu_*.log
