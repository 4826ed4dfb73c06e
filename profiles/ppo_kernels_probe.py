"""Probe: every hand-written kernel of the PPO path (config 3 shapes: 16,384 rows) launched a few times on data from a
real short rollout -- the target of `ncu --set full -k regex:kz_` (profiles/run_r2_ncu.sh) and, run plainly, a table of
CUDA-event times with the bytes each kernel needs."""
import json, os, sys, torch
sys.path.insert(0, ".")
from types import SimpleNamespace
from shogidrl_b200 import rl, nn_ops
from shogidrl_b200.core import ActorCritic, PPOAgent, RolloutBuffer
from shogidrl_b200.training import VecStepManager
from shogidrl_b200 import VecShogiEnv

dev = torch.device("cuda:0")
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
N, T = 16384, int(os.environ.get("T", 24))
cfg = SimpleNamespace(env=SimpleNamespace(device="cuda", seed=1, input_channels=46, num_actions_total=13527, max_moves_per_game=500),
                      training=SimpleNamespace(learning_rate=3e-4, gamma=0.99, lambda_gae=0.95, clip_epsilon=0.2, value_loss_coeff=0.5,
                                               entropy_coef=0.01, ppo_epochs=1, minibatch_size=N, steps_per_epoch=N * T,
                                               total_timesteps=N * T, gradient_clip_max_norm=0.5, normalize_advantages=True,
                                               enable_value_clipping=False, weight_decay=0.0, lr_schedule_type=None,
                                               lr_schedule_step_on="epoch", lr_schedule_kwargs=None))
torch.manual_seed(0)
agent = PPOAgent(ActorCritic(46, 13527), cfg, dev, use_mixed_precision=True)
env = VecShogiEnv(N, 500, dev, seed=1)
buf = RolloutBuffer(T, N, 0.99, 0.95, dev)
drv = VecStepManager(env, agent, buf)
drv.collect(); drv.finish()
torch.cuda.synchronize()
t = T - 1
bm, obs, acts = buf.bitmaps[t], buf.obs[t], buf.actions[t].contiguous()
legal = float(env.legal_count.float().mean())
logits = torch.randn(N, 13536, device=dev).bfloat16()[:, :13527]
rows = torch.randperm(N * T, device=dev)[:N]
flat_bm = buf.bitmaps[:T].reshape(N * T, 448)
flat_obs = buf.obs[:T].reshape(N * T, 46, 9, 9)


def timed(fn, reps=int(os.environ.get("REPS", 20))):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(name, ms, need_bytes, note=""):
    print(f"{name:34s} {ms * 1e3:8.1f} us   needs {need_bytes / 1e6:8.1f} MB -> {need_bytes / ms / 1e6:7.0f} GB/s = "
          f"{need_bytes / ms / 1e6 / peak:.2f} of HBM peak  {note}")


sect = legal * 32
report("kz_sample_bitmap", timed(lambda: rl.sample_masked(logits, bm, seed=1, offset=0)), N * (1792 + sect + 12), f"({legal:.0f} legal/row)")
lg = logits.detach().clone().requires_grad_(True)
report("kz_eval_bitmap_fwd (rows gather)", timed(lambda: rl.evaluate_masked(lg, flat_bm, acts, mask_rows=rows)), N * (1792 + 2 * sect + 24))
lp, en = rl.evaluate_masked(lg, flat_bm, acts, mask_rows=rows)
g1, g2 = torch.ones_like(lp), torch.ones_like(en)
report("kz_eval_bitmap_bwd", timed(lambda: torch.autograd.grad((lp, en), lg, (g1, g2), retain_graph=True)),
       N * (1792 + sect + 27072 + 16), "(dense 27 KB dlogits row)")
w = agent.model.conv.weight.detach().clone().requires_grad_(True)
b = agent.model.conv.bias.detach().clone().requires_grad_(True)
report("kz_obs_conv_fwd (rows gather)", timed(lambda: nn_ops.obs_conv(flat_obs, w, b, relu=True, rows=rows)), N * (14904 + 2592))
y = nn_ops.obs_conv(flat_obs, w, b, relu=True, rows=rows)
dy = torch.randn_like(y)
report("kz_obs_conv_wgrad (+reduce)", timed(lambda: torch.autograd.grad(y, (w, b), dy, retain_graph=True)), N * (14904 + 2 * 2592))
flat_cobs = buf.cobs[:T].reshape(N * T, 40)
report("kz_cobs_conv_fwd (rows gather)", timed(lambda: nn_ops.obs_conv(flat_obs, w, b, relu=True, rows=rows, cobs=flat_cobs)), N * (160 + 2592))
yc = nn_ops.obs_conv(flat_obs, w, b, relu=True, rows=rows, cobs=flat_cobs)
report("kz_cobs_conv_wgrad (+reduce)", timed(lambda: torch.autograd.grad(yc, (w, b), dy, retain_graph=True)), N * (160 + 2 * 2592))
for p in agent.model.parameters():
    p.grad = torch.randn_like(p) * 1e-3
nparam = sum(p.numel() for p in agent.model.parameters())
report("kz_adam_clip_step (3 launches)", timed(lambda: rl.adam_clip_step(agent.optimizer, 0.5)), nparam * 4 * (1 + 7))
report("kz_gae [128 x 16384]", timed(lambda: buf.compute_advantages_and_returns(buf.values[0])), N * T * 17)
a2 = [buf.actions[t].clone(), torch.empty_like(buf.actions[t])]
env.refresh(random_actions=True, next_out=a2[0])
i = [0]


def step_rollout():
    env.step_rollout(a2[i[0] & 1], buf.obs[(i[0] % T) + 1], buf.bitmaps[(i[0] % T) + 1], random_actions=True, next_out=a2[(i[0] + 1) & 1])
    i[0] += 1


report("kz_step_rollout (obs + bitmap)", timed(step_rollout), N * (14904 + 1792 + 7 + 8 + 256 + 528))
