# round 2: 2-GPU runs (NCCL): headline bench with the hardware multi-GPU equality check, U-Net config, trit-plane config
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N exit $?" >> gpurun_out/summary.txt
timeout 900 $TR bench.py --gpus $N --config unet --steps 3 --warmup 3 > gpurun_out/bench_unet_n$N.json 2> gpurun_out/bench_unet_n$N.err; echo "bench unet n$N exit $?" >> gpurun_out/summary.txt
timeout 900 $TR bench.py --gpus $N --config tritplane --batch 2 --steps 20 --warmup 3 > gpurun_out/bench_trit_n$N.json 2> gpurun_out/bench_trit_n$N.err; echo "bench trit n$N exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
for f in bench_n$N bench_unet_n$N bench_trit_n$N; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "parity", json.dumps(d.get("parity"))[:600])
except Exception as e:
    print("$f", "no line:", e)
PY
grep -v "^$" gpurun_out/$f.err | tail -n 3
done
