"""Probe: where does one PPO minibatch update (16,384 samples, default CNN, bf16 autocast) spend its time?"""
import sys, torch
sys.path.insert(0, ".")
from types import SimpleNamespace
from shogidrl_b200.core import ActorCritic, PPOAgent
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda")
import os
torch.backends.cudnn.benchmark = os.environ.get("CUDNN_BENCH", "0") == "1"
cfg = SimpleNamespace(env=SimpleNamespace(device="cuda", seed=1, input_channels=46, num_actions_total=13527, max_moves_per_game=500),
    training=SimpleNamespace(learning_rate=3e-4, gamma=0.99, lambda_gae=0.95, clip_epsilon=0.2, value_loss_coeff=0.5, entropy_coef=0.01,
        ppo_epochs=1, minibatch_size=16384, steps_per_epoch=65536, total_timesteps=1 << 20, gradient_clip_max_norm=0.5,
        normalize_advantages=True, enable_value_clipping=False, weight_decay=0.0, lr_schedule_type=None, lr_schedule_step_on="epoch",
        lr_schedule_kwargs=None), display=SimpleNamespace(display_moves=False, turn_tick=0.0))
agent = PPOAgent(ActorCritic(46, 13527), cfg, dev, use_mixed_precision=True)
B = 65536
class Buf:
    def get_batch(self):
        g = torch.Generator(device="cuda").manual_seed(0)
        mask = torch.rand(B, 13536, device=dev, generator=g) < 0.004
        mask[:, 0] = True
        return {"obs": torch.rand(B, 46, 9, 9, device=dev, generator=g), "actions": torch.zeros(B, dtype=torch.int64, device=dev),
                "log_probs": torch.full((B,), -3.0, device=dev), "values": torch.zeros(B, device=dev),
                "advantages": torch.randn(B, device=dev, generator=g), "returns": torch.randn(B, device=dev, generator=g),
                "legal_masks": mask[:, :13527]}
buf = Buf()
agent.learn(buf)
torch.cuda.synchronize()
import time
t = time.perf_counter(); agent.learn(buf); torch.cuda.synchronize(); dt = time.perf_counter() - t
print(f"learn(): {dt*1e3:.1f} ms for 4 minibatch updates -> {dt*1e3/4:.1f} ms per update")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    agent.learn(buf); torch.cuda.synchronize()
rows = [(e.self_device_time_total / 4, e.count // 4 if e.count >= 4 else e.count, e.key) for e in prof.key_averages() if e.self_device_time_total > 0]
rows.sort(reverse=True)
print("device time per minibatch update (us), launches, kernel")
for t, c, k in rows[:40]:
    print(f"{t:9.1f} {c:4d}  {k[:150]}")
print(f"total {sum(r[0] for r in rows):.0f} us")
