# round 2, call F: first-layer epilogue warps 8 / 12 / 16 (same process, interleaved), eager bench, winattn tests
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
for t in test_gpu_winattn test_gpu_net test_tritplane; do
  timeout 900 python -m pytest tests/$t.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/$t.log 2>&1
  echo "$t exit $? $(tail -n 1 gpurun_out/$t.log)"
done
python tools/ab_inprocess.py first_epi 8 12 16 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_nograph.json 2> gpurun_out/bench_nograph.err; echo "bench nograph exit $?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_nograph.json").read().strip().splitlines()[-1])
print("nograph", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "eager_ms", d.get("eager_ms_per_step"), "winattn", d["roofline_window_attention"]["frac"], d["roofline_window_attention"]["ms"])
PY
