# round 2, call E: uint8 builder with bf16 staging (tests + A/B against fp32 input), SM-partition sweep
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
for t in test_gpu_round2 test_gpu_conv test_gpu_net; do
  timeout 900 python -m pytest tests/$t.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/$t.log 2>&1
  echo "$t exit $? $(tail -n 1 gpurun_out/$t.log)" >> gpurun_out/summary.txt
done
for rep in 1 2; do
for v in u8 f32; do
  timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --input $v > gpurun_out/bench_in${v}_$rep.json 2> gpurun_out/bench_in${v}_$rep.err; echo "bench $v rep $rep exit $?" >> gpurun_out/summary.txt
done
done
for s in 8 12 16 24 28; do
  timeout 600 python bench.py --steps 30 --warmup 5 --side-sms $s --no-cpu-baseline > gpurun_out/bench_side$s.json 2> gpurun_out/bench_side$s.err; echo "bench side $s exit $?" >> gpurun_out/summary.txt
done
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_side0.json 2> gpurun_out/bench_side0.err
cat gpurun_out/summary.txt
for f in bench_inu8_1 bench_inf32_1 bench_inu8_2 bench_inf32_2 bench_side8 bench_side12 bench_side16 bench_side24 bench_side28 bench_side0; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    pl=d["roofline"].get("per_layer_ms_per_step") or {}
    print("$f", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "roofline", round(d["roofline"]["frac"],3), "first", pl.get("kind12_512x768_3->192"), "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("$f", "no line:", e)
PY
done
