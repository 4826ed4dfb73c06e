"""Probe: the fused observation input layer (kz_obs_conv_fwd / kz_obs_conv_wgrad) vs cuDNN under bf16 autocast."""
import json, os, sys, torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from shogidrl_b200 import _native as nv, nn_ops
dev = torch.device("cuda")
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
def timed(fn, reps=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
n = 16384
obs = torch.rand(n, 46, 9, 9, device=dev)
w = (torch.randn(16, 46, 3, 3, device=dev) * 0.05).requires_grad_()
b = torch.zeros(16, device=dev, requires_grad=True)
L = nv.lib(); st = nv.stream_ptr(dev)
y = torch.empty(n, 16, 9, 9, dtype=torch.bfloat16, device=dev)
dy = torch.randn(n, 16, 9, 9, device=dev).bfloat16()
ctas = L.kz_obs_conv_wgrad_ctas(n)
ws = torch.empty(ctas * 16 * 432, device=dev); dw = torch.empty(16, 46, 3, 3, device=dev); db = torch.empty(16, device=dev)
wd, bd = w.detach(), b.detach()
ms = timed(lambda: L.kz_obs_conv_fwd(obs.data_ptr(), None, wd.data_ptr(), bd.data_ptr(), 16, n, 1, y.data_ptr(), st))
by = n * (46 * 81 * 4 + 16 * 81 * 2)
print(f"kz_obs_conv_fwd: {ms*1e3:.1f} us / {n} boards, {by/ms/1e6:.0f} GB/s = {by/ms/1e6/peak:.2f} of measured HBM peak")
ms = timed(lambda: L.kz_obs_conv_wgrad(obs.data_ptr(), None, y.data_ptr(), dy.data_ptr(), 1, 16, n, ws.data_ptr(), ctas,
                                       dw.data_ptr(), db.data_ptr(), st))
by = n * (46 * 81 * 4 + 2 * 16 * 81 * 2)
print(f"kz_obs_conv_wgrad (+reduce): {ms*1e3:.1f} us, {by/ms/1e6:.0f} GB/s = {by/ms/1e6/peak:.2f}")
def cudnn():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        yy = F.relu(F.conv2d(obs, w, b, padding=1))
    yy.backward(dy)
def fused():
    nn_ops.obs_conv(obs, w, b, True).backward(dy)
print(f"fwd+bwd through autograd: cuDNN {timed(cudnn)*1e3:.0f} us, fused {timed(fused)*1e3:.0f} us")
