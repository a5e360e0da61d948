# ncu --set full of the 21 conv / likelihood launches of one eager single-stream step (final code) -> raw CSV
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --side-sms 0 > gpurun_out/b_eager.json 2>/dev/null; echo "plain exit $?"
timeout 1500 ncu --set full --clock-control none -k "regex:conv_tc2|conv_first|k_likelihood" -s 63 -c 21 -o gpurun_out/prof_step -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --side-sms 0 > gpurun_out/ncu_step.log 2>&1
echo "ncu step exit $?"
ncu -i gpurun_out/prof_step.ncu-rep --page raw --csv > gpurun_out/prof_step_raw.csv 2>/dev/null
rm -f gpurun_out/prof_step.ncu-rep
python tools/ncu_summary.py gpurun_out/prof_step_raw.csv | cut -c1-150
