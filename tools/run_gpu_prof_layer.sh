# usage: bash tools/run_gpu_prof_layer.sh <layer> <kernel regex>
mkdir -p gpurun_out
L=${1:-deconv3}; K=${2:-conv_tc2}
timeout 120 python tools/prof_layer.py $L 3 > gpurun_out/prof_layer_$L.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -o gpurun_out/prof_$L -f python tools/prof_layer.py $L 3 > gpurun_out/ncu_$L.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_$L.log
