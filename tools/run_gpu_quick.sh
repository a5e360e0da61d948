# quick loop: the test files named in $TESTS (default: the U-Net tests), output on stdout
mkdir -p gpurun_out
for t in ${TESTS:-test_gpu_unet}; do
  timeout 900 python -m pytest tests/$t.py -q -m gpu --no-header -p no:cacheprovider -s > gpurun_out/$t.log 2>&1
  echo "$t exit $?"
  grep -n "deviations\|passed\|failed\|^E  " gpurun_out/$t.log | cut -c1-1500 | tail -n 30
done
