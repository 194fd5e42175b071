"""Behaviour metrics of ``test_agent``, vectorised over contexts instead of Python loops."""
from __future__ import annotations

import torch


def mode_entropy(mode_encoding: torch.Tensor, successes: torch.Tensor, n_modes: int) -> tuple[torch.Tensor, torch.Tensor]:
    """``pushing_sim.py:140-167``: p(mode | context) over *successful* rollouts, normalised entropy averaged over contexts.

    mode_encoding, successes: [n_contexts, n_traj].  Returns (mode_probs [n_contexts, n_modes], entropy scalar).
    """
    n_traj = mode_encoding.shape[1]
    ok = successes == 1
    probs = torch.stack([((mode_encoding == k) & ok).sum(1).float() / n_traj for k in range(n_modes)], 1)
    probs = probs / (probs.sum(1, keepdim=True) + 1e-12)
    ent = -(probs * torch.log(probs + 1e-12) / torch.log(torch.tensor(float(n_modes)))).sum(1).mean()
    return probs, ent


def avoiding_entropy(mode_encoding: torch.Tensor, successes: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """``avoiding_sim.py:128-134``: 9-bit path code -> integer mode; distribution over the *successful* rollouts,
    entropy in log base 24 (the number of feasible paths).

    mode_encoding: [n_traj, 9] bits, successes: [n_traj].  Returns (mode distribution over observed modes, entropy)."""
    weights = 2 ** torch.arange(9, dtype=torch.float32)
    codes = (mode_encoding.float().cpu() * weights).sum(1)
    codes = codes[successes.cpu() == 1]
    if codes.numel() == 0:
        return torch.zeros(0), torch.tensor(0.0)
    _, counts = torch.unique(codes, return_counts=True)
    probs = counts.float() / counts.sum()
    ent = -(probs * (torch.log(probs) / torch.log(torch.tensor(24.0)))).sum()
    return probs, ent


def mode_kl(mode_encoding: torch.Tensor, successes: torch.Tensor, prior: dict | None):
    """``sorting_sim.py:191-213``: p(mode | context) over successful rollouts for the modes listed in the prior
    ({mode id: probability}, ``<k>_mode_prob.pkl``), contexts without a success dropped, entropy and KL to the prior in log
    base n_mode.  ``prior=None`` (the pkl is not shipped): uniform prior over the modes observed in successful rollouts.

    Returns (mode_probs [n_kept_contexts, n_mode], entropy, KL) as (tensor, float, float)."""
    ok = successes == 1
    if prior is None:
        keys = torch.unique(mode_encoding[ok])
        if keys.numel() == 0:
            return torch.zeros(0, 0), 0.0, 0.0
        prior_p = torch.full((keys.numel(),), 1.0 / keys.numel())
    else:
        keys = torch.tensor(list(prior.keys()), dtype=mode_encoding.dtype)
        prior_p = torch.tensor([float(prior[k]) for k in prior.keys()])
    n_mode, n_traj = keys.numel(), mode_encoding.shape[1]
    probs = torch.stack([((mode_encoding == k) & ok).sum(1).float() / n_traj for k in keys], 1)
    probs = probs / (probs.sum(1, keepdim=True) + 1e-12)
    probs = probs[probs.sum(1) != 0]
    if probs.shape[0] == 0:
        return probs, 0.0, 0.0
    base = torch.log(torch.tensor(float(max(n_mode, 2))))
    entropy = -(probs * torch.log(probs + 1e-12) / base).sum(1).mean()
    log_ = (probs * torch.log(prior_p + 1e-12) / base).sum(1).mean()
    return probs, float(entropy), float(-entropy - log_)
