# round 2, call B: round-2 tests after the fixes, high-model bench, net tests
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
for t in test_gpu_round2 test_gpu_net test_gpu_conv; do
  timeout 900 python -m pytest tests/$t.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/$t.log 2>&1
  echo "$t exit $?" >> gpurun_out/summary.txt
done
timeout 900 python bench.py --config high --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_high.json 2> gpurun_out/bench_high.err; echo "bench high exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_nograph.json 2> gpurun_out/bench_nograph.err; echo "bench nograph exit $?" >> gpurun_out/summary.txt
tail -n 40 gpurun_out/test_gpu_round2.log
tail -n 6 gpurun_out/test_gpu_conv.log gpurun_out/test_gpu_net.log
cat gpurun_out/summary.txt
for f in bench_high bench_nograph; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "eager_ms", d.get("eager_ms_per_step"), "roofline", round(d["roofline"]["frac"],3))
    print(json.dumps(d["roofline"].get("per_layer_ms_per_step")))
    print(json.dumps(d["roofline"].get("per_layer_tflops")))
except Exception as e:
    print("$f", "no line:", e)
PY
done
grep -v "^$" gpurun_out/bench_high.err | tail -n 4
