mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_entropy.py -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/test_gpu_entropy.log 2>&1; echo "entropy exit $?" >> gpurun_out/summary.txt
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $BENCH > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launch.log 2>&1
echo "launchlist exit $?" >> gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 14 -c 14 -o gpurun_out/prof_conv -f $BENCH > gpurun_out/ncu_conv.log 2>&1
echo "ncu conv exit $?" >> gpurun_out/summary.txt
timeout 120 python tools/prof_likelihood.py 5 > gpurun_out/lik_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_likelihood -s 3 -c 1 -o gpurun_out/prof_lik -f python tools/prof_likelihood.py 5 > gpurun_out/ncu_lik.log 2>&1
echo "ncu lik exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt gpurun_out/lik_plain.log; tail -3 gpurun_out/test_gpu_entropy.log; ls -la gpurun_out
