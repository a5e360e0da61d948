# ncu capture of the rANS kernels at the bench size (one encode + one decode at 2048 symbols per stream)
# usage: bash tools/prof_rans_ncu.sh [kernel regex, default k_rans]
K=${1:-k_rans}
mkdir -p gpurun_out
cat > /tmp/rans_once.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from ldic_b200 import ops
torch.cuda.set_device(0)
g = torch.Generator(device="cuda").manual_seed(1)
B, h, w, N, M, Cp = 16, 32, 48, 192, 16, 192
Cc, P = N - M, B * h * w
ctx = torch.empty(P, 2 * Cp, device="cuda")
ctx[:, :Cp] = torch.randn(P, Cp, device="cuda", generator=g) * 3
ctx[:, Cp:] = torch.randn(P, Cp, device="cuda", generator=g) * 0.8 + 0.2
y = torch.zeros(B, h, w, N, device="cuda")
y[..., M:] = ctx[:, :Cc].reshape(B, h, w, Cc) + torch.exp(ctx[:, Cp:Cp + Cc]).reshape(B, h, w, Cc) * torch.randn(B, h, w, Cc, device="cuda", generator=g)
kw = dict(mu=ctx, mu_mode=2, mu_rs=2 * Cp, sigma=ctx, sigma_mode=2, sigma_rs=2 * Cp, sigma_off=Cp, sigma_is_log=True)
out = torch.empty(B, h, w, Cc, device="cuda")
for _ in range(2):
    enc = ops.rans_encode_rows(y, P, Cc, h * w, v_rs=N, v_off=M, **kw)
    ops.rans_decode_rows(enc, P, Cc, h * w, out, v_hat_rs=Cc, **kw)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:$K -c 18 -o gpurun_out/rans_full -f python /tmp/rans_once.py > gpurun_out/rans_ncu.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/rans_full.ncu-rep --page raw --csv > gpurun_out/rans_full_raw.csv 2> /dev/null
ncu -i gpurun_out/rans_full.ncu-rep --page source --csv --kernel-name regex:k_rans_enc_streams > gpurun_out/rans_enc_source.csv 2> /dev/null
ncu -i gpurun_out/rans_full.ncu-rep --page source --csv --kernel-name regex:k_rans_dec_streams > gpurun_out/rans_dec_source.csv 2> /dev/null
rm -f gpurun_out/rans_full.ncu-rep
ls -la gpurun_out | tail -n 6
