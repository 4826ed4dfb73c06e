"""Probe: device time of the agent/buffer kernels through the C ABI with preallocated buffers
(kz_sample_masked, kz_eval_masked_fwd/bwd, kz_gae), against the measured HBM peak."""
import json, os, sys, torch
sys.path.insert(0, ".")
from shogidrl_b200 import _native as nv
dev = torch.device("cuda")
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
L = nv.lib(); st = nv.stream_ptr(dev)
def timed(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
n, A, LD = 16384, 13527, 13536
g = torch.Generator(device="cuda").manual_seed(0)
mask = (torch.rand(n, LD, device=dev, generator=g) < 0.004).to(torch.uint8)
mask[:, 0] = 1
legal = float(mask[:, :A].sum()) / n
acts = torch.zeros(n, dtype=torch.int64, device=dev)
logp = torch.empty(n, device=dev); ent = torch.empty(n, device=dev); saved = torch.empty(n, 4, device=dev)
ones = torch.ones(n, device=dev)
for dtype, nb in ((torch.bfloat16, 2), (torch.float32, 4)):
    logits = torch.randn(n, LD, device=dev, generator=g).to(dtype)
    grad = torch.empty_like(logits)
    bf = int(dtype == torch.bfloat16)
    need = n * (LD + legal * 32)   # mask row + one 32-byte sector per legal logit
    ms = timed(lambda: L.kz_sample_masked(logits.data_ptr(), bf, LD, mask.data_ptr(), LD, n, 1, 0, acts.data_ptr(), 1,
                                          logp.data_ptr(), None, 0, st))
    print(f"kz_sample_masked {dtype} ({legal:.0f} legal/row): {ms*1e3:.1f} us / {n} rows; dense-equivalent "
          f"{n*A*(nb+1)/ms/1e6:.0f} GB/s, bytes needed {need/ms/1e6:.0f} GB/s = {need/ms/1e6/peak:.2f} of HBM peak")
    ms = timed(lambda: L.kz_eval_masked_fwd(logits.data_ptr(), bf, LD, mask.data_ptr(), LD, None, acts.data_ptr(), n,
                                            logp.data_ptr(), ent.data_ptr(), saved.data_ptr(), st))
    print(f"kz_eval_masked_fwd {dtype}: {ms*1e3:.1f} us; bytes needed {need/ms/1e6:.0f} GB/s = {need/ms/1e6/peak:.2f}")
    ms = timed(lambda: L.kz_eval_masked_bwd(logits.data_ptr(), bf, LD, mask.data_ptr(), LD, None, acts.data_ptr(), n,
                                            ones.data_ptr(), ones.data_ptr(), saved.data_ptr(), grad.data_ptr(), LD, st))
    needb = need + n * A * nb
    print(f"kz_eval_masked_bwd {dtype}: {ms*1e3:.1f} us; bytes needed (dense dlogits write) {needb/ms/1e6:.0f} GB/s = "
          f"{needb/ms/1e6/peak:.2f}")
for T, N in ((128, 16384), (128, 65536), (2048, 1)):
    r, v = torch.randn(T, N, device=dev), torch.randn(T, N, device=dev)
    d = (torch.rand(T, N, device=dev) < 0.01).to(torch.uint8); lv = torch.randn(N, device=dev)
    adv, ret = torch.empty_like(r), torch.empty_like(r)
    for name in ("kz_gae", "kz_gae_exact"):
        ms = timed(lambda: getattr(L, name)(r.data_ptr(), v.data_ptr(), d.data_ptr(), lv.data_ptr(), T, N, 0.99, 0.99 * 0.95,
                                            adv.data_ptr(), ret.data_ptr(), st))
        print(f"{name} [T={T}, N={N}]: {ms*1e3:.1f} us, {T*N*17/ms/1e6:.0f} GB/s = {T*N*17/ms/1e6/peak:.2f}")
