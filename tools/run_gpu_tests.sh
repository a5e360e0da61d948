mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
for t in test_gpu_entropy test_gpu_conv test_gpu_net; do
  timeout 300 python -m pytest tests/$t.py -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/$t.log 2>&1
  echo "$t exit $?" >> gpurun_out/summary.txt
done
tail -n 30 gpurun_out/test_gpu_entropy.log gpurun_out/test_gpu_conv.log gpurun_out/test_gpu_net.log
cat gpurun_out/summary.txt
