mkdir -p gpurun_out
for t in test_gpu_conv test_gpu_net; do
  timeout 300 python -m pytest tests/$t.py -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/$t.log 2>&1
  echo "$t exit $?" >> gpurun_out/summary.txt
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 4 gpurun_out/test_gpu_conv.log gpurun_out/test_gpu_net.log; tail -n 5 gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json"))
print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "conv_ms", round(d["roofline"]["conv_ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 3), "lik", round(d["roofline_likelihood"]["frac"], 3))
print(d["roofline"]["per_layer_tflops"]); print(d["parity"])
PY
