mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_entropy.py tests/test_gpu_net.py -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/test_entropy_net.log 2>&1; echo "tests exit $?"
tail -n 5 gpurun_out/test_entropy_net.log
timeout 120 python tools/prof_likelihood.py 8
