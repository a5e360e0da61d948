# round 2, call C: U-Net family tests + bench, round-2 tests
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
for t in test_gpu_unet test_gpu_round2 test_gpu_winattn; do
  timeout 900 python -m pytest tests/$t.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/$t.log 2>&1
  echo "$t exit $?" >> gpurun_out/summary.txt
done
timeout 900 python bench.py --config unet --batch 8 --steps 5 --warmup 3 > gpurun_out/bench_unet_b8.json 2> gpurun_out/bench_unet_b8.err; echo "bench unet b8 exit $?" >> gpurun_out/summary.txt
tail -n 40 gpurun_out/test_gpu_unet.log
tail -n 12 gpurun_out/test_gpu_round2.log gpurun_out/test_gpu_winattn.log
cat gpurun_out/summary.txt
cat gpurun_out/bench_unet_b8.json; grep -v "^$" gpurun_out/bench_unet_b8.err | tail -n 5
