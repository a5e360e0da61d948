mkdir -p gpurun_out
for t in test_gpu_entropy test_gpu_conv test_gpu_net; do
  timeout 300 python -m pytest tests/$t.py -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/$t.log 2>&1
  echo "$t exit $?" >> gpurun_out/summary.txt
done
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench_ref exit $?" >> gpurun_out/summary.txt
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
tail -n 5 gpurun_out/test_gpu_*.log gpurun_out/smoke.log
cat gpurun_out/summary.txt gpurun_out/bench.json gpurun_out/bench_ref.json; tail -5 gpurun_out/bench.err
