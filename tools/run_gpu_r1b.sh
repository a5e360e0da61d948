# tests + A/B of the halo B-slot grouping + launch list + full ncu capture of one step
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
for t in test_gpu_entropy test_gpu_conv test_gpu_net; do
  timeout 300 python -m pytest tests/$t.py -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/$t.log 2>&1
  echo "$t exit $?" >> gpurun_out/summary.txt
done
for cfg in "LDIC_HALO_G=4" "LDIC_HALO_G=1"; do
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err
  echo "bench $cfg exit $?" >> gpurun_out/summary.txt
done
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launch.log 2>&1
echo "launchlist exit $?" >> gpurun_out/summary.txt
timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:conv_tc_kernel|conv_halo_kernel|k_im2col|k_likelihood|k_syntax" -s 69 -c 23 -o gpurun_out/prof_step -f $BENCH > gpurun_out/ncu_step.log 2>&1
echo "ncu step exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 4 gpurun_out/test_gpu_*.log
python - <<'PY'
import json
for c in ("LDIC_HALO_G=4", "LDIC_HALO_G=1"):
    try:
        d = json.load(open(f"gpurun_out/bench_{c}.json"))
        print(c, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "conv_ms", round(d["roofline"]["conv_ms_per_step"], 3))
        print(d["roofline"]["per_layer_ms_per_step"])
    except Exception as e:
        print(c, "failed", e)
PY
