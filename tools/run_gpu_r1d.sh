mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
for t in test_gpu_net test_gpu_conv test_gpu_entropy; do
  timeout 300 python -m pytest tests/$t.py -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/$t.log 2>&1
  echo "$t exit $?" >> gpurun_out/summary.txt
done
for cfg in "LDIC_X=0" "LDIC_SYNTAX_FUSED=0"; do
  env $cfg timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err
  echo "bench $cfg exit $?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt; tail -n 25 gpurun_out/test_gpu_net.log; tail -n 4 gpurun_out/test_gpu_conv.log gpurun_out/test_gpu_entropy.log
python - <<'PY'
import json
for c in ("LDIC_X=0", "LDIC_SYNTAX_FUSED=0"):
    try:
        d = json.load(open(f"gpurun_out/bench_{c}.json"))
        print(c, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "conv_ms", round(d["roofline"]["conv_ms_per_step"], 3), d["parity"], d["gpu_launches"], d["clocks"])
    except Exception as e:
        print(c, "failed", e)
PY
tail -n 5 gpurun_out/bench_*.err
