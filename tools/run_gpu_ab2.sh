mkdir -p gpurun_out
for rep in 1 2; do
for cfg in "LDIC_SYNTAX_FUSED=0" "LDIC_SYNTAX_FUSED=1"; do
  env $cfg timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
  python - "$cfg" <<'PY'
import json, sys
d = json.load(open("gpurun_out/bench_ab.json"))
print(sys.argv[1], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "conv_ms", round(d["roofline"]["conv_ms_per_step"], 3), d["parity"], d["gpu_launches"], d["clocks"])
PY
done; done
