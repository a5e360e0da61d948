python tools/exp_timing.py 2>&1 | grep -E "timing|conv2|gs4" | head -6
bash tools/run_gpu_quick.sh
