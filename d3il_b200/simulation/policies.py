"""Synthetic stand-ins for the reference's policies, for benchmarks and plumbing tests only.

The agents themselves are out of scope (reused unchanged through ``agent_adapter``); trained
checkpoints and datasets are not available offline.  These modules have the *shapes* of the reference policies so the
rollout loop does the same amount of policy work per env step:

* ``ResidualMLP`` — Linear(obs,H) -> L/2 pre-activation residual blocks (Mish) -> Linear(H,act)
  (``agents/models/common/mlp.py:9-46,114-190``).
* ``SyntheticBCPolicy`` — BC-MLP ``predict`` on a batch: scale -> net -> clamp -> unscale (``agents/bc_agent.py:241-271``),
  identity scaler, action bounds +-0.01 (the env's ``action_space``).
* ``SyntheticDDPMPolicy`` — DDPM-MLP: T reverse-diffusion steps of an epsilon-predicting ResidualMLP over
  [x, time-embedding, state], cosine schedule, clipped x0 (``agents/models/diffusion/gc_diffusion.py:100-216``,
  ``diffusion_models.py:20-115``); Sorting-4 script sizes H=256, L=8, T=4, t_dim=8.

Random-init weights (seeded); both expose ``predict_batch(obs[N, obs_dim]) -> [N, act_dim]`` and ``reset()``.
"""
from __future__ import annotations

import math

import torch
from torch import nn


class _ResBlock(nn.Module):
    def __init__(self, width: int):
        super().__init__()
        self.a, self.b, self.f = nn.Linear(width, width), nn.Linear(width, width), nn.Mish()

    def forward(self, x):
        return x + self.b(self.f(self.a(self.f(x))))


class ResidualMLP(nn.Module):
    def __init__(self, n_in: int, width: int, n_hidden_layers: int, n_out: int):
        super().__init__()
        if n_hidden_layers % 2:
            raise ValueError("the reference network needs an even number of hidden layers")
        self.net = nn.Sequential(nn.Linear(n_in, width), *[_ResBlock(width) for _ in range(n_hidden_layers // 2)], nn.Linear(width, n_out))

    def forward(self, x):
        return self.net(x)


class SyntheticBCPolicy:
    graph_safe = True          # pure tensor code: the rollout body can be captured in a CUDA graph

    def __init__(self, obs_dim: int, act_dim: int, width: int = 128, n_hidden_layers: int = 6, bound: float = 0.01, device="cuda", seed: int = 0):
        g = torch.random.fork_rng(devices=[])
        with g:
            torch.manual_seed(seed)
            self.model = ResidualMLP(obs_dim, width, n_hidden_layers, act_dim).to(device).eval()
        self.bound, self.device = bound, torch.device(device)

    def reset(self):
        pass

    @torch.no_grad()
    def predict_batch(self, obs: torch.Tensor) -> torch.Tensor:
        return self.model(obs.to(self.device, torch.float32)).clamp_(-1.0, 1.0) * self.bound

    def predict(self, obs):            # single-sample API of the reference agents: np[obs_dim] -> np[1, act_dim]
        import numpy as np
        return self.predict_batch(torch.as_tensor(np.asarray(obs), dtype=torch.float32)[None]).cpu().numpy()


class SyntheticDDPMPolicy:
    graph_safe = True          # pure tensor code (torch.randn on the default CUDA generator is graph-capturable)

    def __init__(self, obs_dim: int, act_dim: int, width: int = 256, n_hidden_layers: int = 8, n_timesteps: int = 4, t_dim: int = 8,
                 bound: float = 0.01, device="cuda", seed: int = 0):
        with torch.random.fork_rng(devices=[]):
            torch.manual_seed(seed)
            self.eps_net = ResidualMLP(act_dim + t_dim + obs_dim, width, n_hidden_layers, act_dim).to(device).eval()
            self.t_mlp = nn.Sequential(nn.Linear(t_dim, 2 * t_dim), nn.Mish(), nn.Linear(2 * t_dim, t_dim)).to(device).eval()
        self.device, self.bound, self.T, self.t_dim, self.act_dim = torch.device(device), bound, n_timesteps, t_dim, act_dim
        # cosine schedule (Nichol & Dhariwal), s = 0.008, betas clipped at 0.999
        k = torch.linspace(0, n_timesteps + 1, n_timesteps + 1, dtype=torch.float64)
        ac = torch.cos(((k / (n_timesteps + 1)) + 0.008) / 1.008 * math.pi / 2) ** 2
        ac = ac / ac[0]
        beta = (1 - ac[1:] / ac[:-1]).clamp(0, 0.999)
        alpha = 1 - beta
        abar = torch.cumprod(alpha, 0)
        abar_prev = torch.cat([torch.ones(1, dtype=torch.float64), abar[:-1]])
        f = lambda v: v.to(torch.float32).to(device)      # noqa: E731
        self.c_x0_from_xt, self.c_x0_from_eps = f(torch.sqrt(1 / abar)), f(torch.sqrt(1 / abar - 1))
        self.c_mean_x0, self.c_mean_xt = f(beta * torch.sqrt(abar_prev) / (1 - abar)), f((1 - abar_prev) * torch.sqrt(alpha) / (1 - abar))
        self.sigma = f(torch.sqrt((beta * (1 - abar_prev) / (1 - abar)).clamp(min=1e-20)))
        half = t_dim // 2
        self.freq = torch.exp(torch.arange(half, device=device) * -(math.log(10000.0) / max(half - 1, 1)))
        # noise comes from the DEFAULT CUDA generator (seed it with torch.manual_seed): that one is registered with CUDA graphs, a private
        # torch.Generator is not ("Attempt to increase offset for a CUDA generator not in capture mode")

    def reset(self):
        pass

    @torch.no_grad()
    def predict_batch(self, obs: torch.Tensor) -> torch.Tensor:
        s = obs.to(self.device, torch.float32)
        n = s.shape[0]
        x = torch.randn(n, self.act_dim, device=self.device)
        for t in range(self.T - 1, -1, -1):
            ang = t * self.freq
            temb = self.t_mlp(torch.cat([ang.sin(), ang.cos()])[None]).expand(n, -1)
            eps = self.eps_net(torch.cat([x, temb, s], 1))
            x0 = (self.c_x0_from_xt[t] * x - self.c_x0_from_eps[t] * eps).clamp_(-1.0, 1.0)
            x = self.c_mean_x0[t] * x0 + self.c_mean_xt[t] * x
            if t > 0:
                x = x + self.sigma[t] * torch.randn(n, self.act_dim, device=self.device)
        return x.clamp_(-1.0, 1.0) * self.bound

    def predict(self, obs):
        import numpy as np
        return self.predict_batch(torch.as_tensor(np.asarray(obs), dtype=torch.float32)[None]).cpu().numpy()
