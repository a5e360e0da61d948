# round 2, call D: first-layer 12-warp epilogue A/B on one box, trit-plane after telescoping, tests
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
for t in test_gpu_conv test_gpu_net test_tritplane test_gpu_unet; do
  timeout 900 python -m pytest tests/$t.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/$t.log 2>&1
  echo "$t exit $? $(tail -n 1 gpurun_out/$t.log)" >> gpurun_out/summary.txt
done
for rep in 1 2; do
for v in 0 8; do
  LDIC_FIRST_EPI=$v timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_epi${v}_$rep.json 2> gpurun_out/bench_epi${v}_$rep.err; echo "bench epi$v rep $rep exit $?" >> gpurun_out/summary.txt
done
done
timeout 300 python bench.py --config tritplane --steps 20 --warmup 3 > gpurun_out/bench_trit.json 2> gpurun_out/bench_trit.err; echo "bench trit exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
for f in bench_epi0_1 bench_epi8_1 bench_epi0_2 bench_epi8_2 bench_trit; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    pl=d["roofline"].get("per_layer_ms_per_step") or {}
    print("$f", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "roofline", round(d["roofline"]["frac"],3), "first", pl.get("kind12_512x768_3->192"), "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("$f", "no line:", e)
PY
done
