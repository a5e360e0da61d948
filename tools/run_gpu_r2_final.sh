# round-2 evidence: all GPU tests, smoke, bench lines (ours + reference arm, sustained, high, trit-plane, U-Net),
# launch list, full ncu capture of one eager step (N=192) and of the wide / trit-plane kernels
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest -m gpu exit $? $(tail -n 1 gpurun_out/pytest_gpu.log)" >> gpurun_out/summary.txt
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $? $(tail -n 1 gpurun_out/smoke.log)" >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench_ref exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 2000 --warmup 20 --no-cpu-baseline > gpurun_out/bench_sustained.json 2> gpurun_out/bench_sustained.err; echo "bench sustained exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_nograph.json 2> gpurun_out/bench_nograph.err; echo "bench nograph exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --config high --steps 10 --warmup 3 > gpurun_out/bench_high.json 2> gpurun_out/bench_high.err; echo "bench high exit $?" >> gpurun_out/summary.txt
timeout 300 python bench.py --config tritplane --steps 20 --warmup 3 > gpurun_out/bench_trit.json 2> gpurun_out/bench_trit.err; echo "bench trit exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --config unet --batch 16 --steps 5 --warmup 3 > gpurun_out/bench_unet_b16.json 2> gpurun_out/bench_unet_b16.err; echo "bench unet b16 exit $?" >> gpurun_out/summary.txt
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launch.log 2>&1
echo "launchlist exit $?" >> gpurun_out/summary.txt
# one whole step, launches issued eagerly on a single stream (21 matching launches per step)
timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:conv_tc2|conv_first|k_likelihood" -s 63 -c 21 -o gpurun_out/prof_step -f $BENCH --no-graph --side-sms 0 > gpurun_out/ncu_step.log 2>&1
echo "ncu step exit $?" >> gpurun_out/summary.txt
timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:conv_wide|conv_tc2" -s 57 -c 19 -o gpurun_out/prof_high -f python bench.py --config high --steps 2 --warmup 3 --no-cpu-baseline --no-graph --side-sms 0 > gpurun_out/ncu_high.log 2>&1
echo "ncu high exit $?" >> gpurun_out/summary.txt
timeout 120 python tools/prof_tritplane.py 5 > gpurun_out/trit_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tritplane -s 3 -c 1 -o gpurun_out/prof_trit -f python tools/prof_tritplane.py 5 > gpurun_out/ncu_trit.log 2>&1
echo "ncu trit exit $?" >> gpurun_out/summary.txt
timeout 120 python tools/prof_winattn.py 5 > gpurun_out/winattn_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_window_attention -s 3 -c 1 -o gpurun_out/prof_winattn -f python tools/prof_winattn.py 5 > gpurun_out/ncu_winattn.log 2>&1
echo "ncu winattn exit $?" >> gpurun_out/summary.txt
timeout 120 python tools/prof_likelihood.py 5 > gpurun_out/lik_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_likelihood -s 3 -c 1 -o gpurun_out/prof_lik -f python tools/prof_likelihood.py 5 > gpurun_out/ncu_lik.log 2>&1
echo "ncu lik exit $?" >> gpurun_out/summary.txt
# reports -> raw CSV pages here (the .ncu-rep files of several full captures exceed what travels back), keep the small ones
for r in prof_step prof_high prof_trit prof_lik prof_winattn; do
  if [ -f gpurun_out/$r.ncu-rep ]; then ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null; fi
done
ncu -i gpurun_out/prof_step.ncu-rep --page source --csv --kernel-name regex:conv_first -c 1 > gpurun_out/src_first.csv 2>/dev/null
rm -f gpurun_out/prof_step.ncu-rep gpurun_out/prof_high.ncu-rep
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
cat gpurun_out/summary.txt; tail -n 3 gpurun_out/pytest_gpu.log gpurun_out/trit_plain.log gpurun_out/lik_plain.log
for f in bench bench_sustained bench_nograph bench_high bench_trit bench_unet_b16 bench_ref; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],2), d["unit"], "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "eager_ms", d.get("eager_ms_per_step"), "roofline", (d.get("roofline") or {}).get("frac"), "clk", (d.get("clocks") or {}).get("sm_mhz"))
except Exception as e:
    print("$f", "no line:", e)
PY
done
