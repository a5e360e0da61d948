# round-end evidence: tests, bench lines (ours + reference arm), launch list, full ncu capture of one step, likelihood kernel capture
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
for t in test_gpu_entropy test_gpu_conv test_gpu_net test_gpu_winattn test_tritplane; do
  timeout 300 python -m pytest tests/$t.py -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/$t.log 2>&1
  echo "$t exit $?" >> gpurun_out/summary.txt
done
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench_ref exit $?" >> gpurun_out/summary.txt
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launch.log 2>&1
echo "launchlist exit $?" >> gpurun_out/summary.txt
# one whole step, launches issued eagerly (21 matching launches per step: 1 conv_first, 16 conv_tc2, 1 conv_halo2, 3 likelihood)
timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:conv_tc|conv_halo|conv_first|k_likelihood" -s 63 -c 21 -o gpurun_out/prof_step -f $BENCH --no-graph > gpurun_out/ncu_step.log 2>&1
echo "ncu step exit $?" >> gpurun_out/summary.txt
timeout 120 python tools/prof_likelihood.py 5 > gpurun_out/lik_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_likelihood -s 3 -c 1 -o gpurun_out/prof_lik -f python tools/prof_likelihood.py 5 > gpurun_out/ncu_lik.log 2>&1
echo "ncu lik exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt gpurun_out/bench.json gpurun_out/bench_ref.json gpurun_out/lik_plain.log; tail -3 gpurun_out/smoke.log
