# round 2: scaling on one 8-GPU box: headline bench at N = 1, 2, 4, 8 (uint8 host buffers, hardware multi-GPU equality
# check at N > 1), U-Net configs[2] at N = 2 / 4 / 8 (global batch 64) and configs[3] at N = 8 (1280x2048, global batch 32),
# trit-plane configs[4] at N = 8 (2 images per GPU)
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
TR() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29533 "${@:2}"; }
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err; echo "n1 exit $?" >> gpurun_out/summary.txt
for n in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; echo "n$n exit $?" >> gpurun_out/summary.txt
done
for n in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --config unet --steps 3 --warmup 3 > gpurun_out/unet_c2_n$n.json 2> gpurun_out/unet_c2_n$n.err; echo "unet c2 n$n exit $?" >> gpurun_out/summary.txt
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --config unet --crop 1280x2048 --steps 3 --warmup 3 > gpurun_out/unet_c3_n8.json 2> gpurun_out/unet_c3_n8.err; echo "unet c3 n8 exit $?" >> gpurun_out/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --config tritplane --batch 2 --steps 20 --warmup 3 > gpurun_out/trit_c4_n8.json 2> gpurun_out/trit_c4_n8.err; echo "trit c4 n8 exit $?" >> gpurun_out/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --config high --steps 10 --warmup 3 > gpurun_out/high_n8.json 2> gpurun_out/high_n8.err; echo "high n8 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
for f in scale_n1 scale_n2 scale_n4 scale_n8 unet_c2_n2 unet_c2_n4 unet_c2_n8 unet_c3_n8 trit_c4_n8 high_n8; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    mg=(d.get("parity") or {}).get("multi_gpu_equal")
    print("$f", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "multi_gpu_equal", (mg or {}).get("equal"), "clk", (d.get("clocks") or {}).get("sm_mhz"))
except Exception as e:
    print("$f", "no line:", e)
PY
grep -v "^$" gpurun_out/$f.err | grep -iv "OMP_NUM_THREADS\|^\*\*\*\|NCCL version" | tail -n 2
done
