# round 2, final code on one 8-GPU box: headline bench, U-Net configs[2] / configs[3], codec (forward + rANS) at N = 8
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
TR() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 "$@"; }
TR --steps 20 --warmup 3 > gpurun_out/final_n8.json 2> gpurun_out/final_n8.err; echo "net n8 exit $?" >> gpurun_out/summary.txt
TR --config codec --steps 20 --warmup 3 > gpurun_out/codec_n8.json 2> gpurun_out/codec_n8.err; echo "codec n8 exit $?" >> gpurun_out/summary.txt
TR --config unet --steps 3 --warmup 3 > gpurun_out/unet_c2_final_n8.json 2> gpurun_out/unet_c2_final_n8.err; echo "unet c2 n8 exit $?" >> gpurun_out/summary.txt
TR --config unet --crop 1280x2048 --steps 3 --warmup 3 > gpurun_out/unet_c3_final_n8.json 2> gpurun_out/unet_c3_final_n8.err; echo "unet c3 n8 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
for f in final_n8 codec_n8 unet_c2_final_n8 unet_c3_final_n8; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    mg=(d.get("parity") or {}).get("multi_gpu_equal")
    print("$f", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "multi_gpu_equal", (mg or {}).get("equal") if isinstance(mg, dict) else mg, "clk", (d.get("clocks") or {}).get("sm_mhz"))
except Exception as e:
    print("$f", "no line:", e)
PY
grep -v "^$" gpurun_out/$f.err | grep -iv "OMP_NUM_THREADS\|^\*\*\*\|NCCL version" | tail -n 2
done
