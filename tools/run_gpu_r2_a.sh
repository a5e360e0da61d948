# round 2, call A: existing suites after the refactor, the new round-2 tests, smoke, short bench lines
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
for t in test_gpu_round2 test_gpu_entropy test_gpu_conv test_gpu_net test_gpu_winattn test_tritplane; do
  timeout 600 python -m pytest tests/$t.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/$t.log 2>&1
  echo "$t exit $?" >> gpurun_out/summary.txt
done
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 --input f32 --no-cpu-baseline > gpurun_out/bench_f32.json 2> gpurun_out/bench_f32.err; echo "bench f32 exit $?" >> gpurun_out/summary.txt
for s in 20 28 36; do
  timeout 600 python bench.py --steps 20 --warmup 3 --side-sms $s --no-cpu-baseline > gpurun_out/bench_side$s.json 2> gpurun_out/bench_side$s.err; echo "bench side $s exit $?" >> gpurun_out/summary.txt
done
timeout 900 python bench.py --config high --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_high.json 2> gpurun_out/bench_high.err; echo "bench high exit $?" >> gpurun_out/summary.txt
timeout 300 python bench.py --config tritplane --steps 20 --warmup 3 > gpurun_out/bench_trit.json 2> gpurun_out/bench_trit.err; echo "bench trit exit $?" >> gpurun_out/summary.txt
tail -n 25 gpurun_out/test_gpu_round2.log
tail -n 6 gpurun_out/test_gpu_entropy.log gpurun_out/test_gpu_conv.log gpurun_out/test_gpu_net.log gpurun_out/test_gpu_winattn.log gpurun_out/test_tritplane.log gpurun_out/smoke.log
cat gpurun_out/summary.txt
for f in bench bench_f32 bench_side20 bench_side28 bench_side36 bench_high bench_trit; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "eager_ms", d.get("eager_ms_per_step"), "roofline", round(d["roofline"]["frac"],3))
except Exception as e:
    print("$f", "no line:", e)
PY
done
tail -3 gpurun_out/bench.err gpurun_out/bench_high.err
