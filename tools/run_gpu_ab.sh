for cfg in "LDIC_HALO=1" "LDIC_HALO=0"; do
  echo "== $cfg"
  env $cfg python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b.json 2>/dev/null
  python - <<'PY'
import json
d = json.load(open("gpurun_out/b.json"))
print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "conv_ms", round(d["roofline"]["conv_ms_per_step"], 3))
print({k: v for k, v in d["roofline"]["per_layer_tflops"].items() if k.startswith(("kind4", "kind5", "kind6", "kind7", "kind8", "kind2"))})
PY
done
