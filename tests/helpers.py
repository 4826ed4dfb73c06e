"""Shared test helpers: a duck-typed stand-in for the reference's pydantic AppConfig (config_schema.py)."""
from types import SimpleNamespace


def make_config(device="cuda", **training):
    tr = dict(learning_rate=3e-4, gamma=0.99, lambda_gae=0.95, clip_epsilon=0.2, value_loss_coeff=0.5,
              entropy_coef=0.01, ppo_epochs=2, minibatch_size=64, steps_per_epoch=128, total_timesteps=1024,
              gradient_clip_max_norm=0.5, normalize_advantages=True, enable_value_clipping=False, weight_decay=0.0,
              lr_schedule_type=None, lr_schedule_step_on="epoch", lr_schedule_kwargs=None)
    tr.update(training)
    return SimpleNamespace(
        env=SimpleNamespace(device=device, seed=42, input_channels=46, num_actions_total=13527, max_moves_per_game=500),
        training=SimpleNamespace(**tr),
        display=SimpleNamespace(display_moves=False, turn_tick=0.0))
