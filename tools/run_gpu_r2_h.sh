# all GPU tests + the bench lines of the final code (no ncu)
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest -m gpu exit $? $(tail -n 1 gpurun_out/pytest_gpu.log)" >> gpurun_out/summary.txt
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $? $(tail -n 1 gpurun_out/smoke.log)" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_nograph.json 2> gpurun_out/bench_nograph.err; echo "bench nograph exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --config unet --batch 16 --steps 5 --warmup 3 > gpurun_out/bench_unet_b16.json 2> gpurun_out/bench_unet_b16.err; echo "bench unet b16 exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --config codec --steps 10 --warmup 3 > gpurun_out/bench_codec.json 2> gpurun_out/bench_codec.err; echo "bench codec exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 3 gpurun_out/pytest_gpu.log
for f in bench bench_nograph bench_unet_b16 bench_codec; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],2), d["unit"], "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "eager_ms", d.get("eager_ms_per_step"), "roofline", (d.get("roofline") or {}).get("frac"), "winattn", (d.get("roofline_window_attention") or {}).get("frac"), d.get("share_of_step"))
except Exception as e:
    print("$f", "no line:", e)
PY
done
