mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unet.py -q -m gpu --no-header -p no:cacheprovider -s > gpurun_out/test_gpu_unet.log 2>&1
echo "test_gpu_unet exit $?"; grep -n "deviations\|passed\|failed\|^E  " gpurun_out/test_gpu_unet.log | cut -c1-700 | tail -n 20
timeout 900 python bench.py --config unet --batch 16 --steps 5 --warmup 3 > gpurun_out/bench_unet_b16.json 2> gpurun_out/bench_unet_b16.err; echo "bench unet exit $?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_unet_b16.json").read().strip().splitlines()[-1])
print("unet b16", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), d["share_of_step"], d["gpu_launches"], d["parity"])
PY
python tools/prof_unet_torch.py 16 2>&1 | tail -48
