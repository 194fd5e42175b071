"""Behaviour metrics of ``test_agent``, vectorised over contexts instead of Python loops.

Every function is plain tensor arithmetic on whatever device its inputs live on: fed with the CUDA result rows of a rollout
they run on the GPU (one small kernel per line, no host round trip); the ``*_Sim`` classes move only the final scalars."""
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
    ent = -(probs * torch.log(probs + 1e-12) / torch.log(torch.tensor(float(n_modes), device=probs.device))).sum(1).mean()
    return probs, ent


def avoiding_entropy(mode_encoding: torch.Tensor, successes: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """``avoiding_sim.py:128-134``: 9-bit path code -> integer mode; distribution over the *successful* rollouts,
    entropy in log base 24 (the number of feasible paths).

    mode_encoding: [n_traj, 9] bits, successes: [n_traj].  Returns (mode distribution over observed modes, entropy)."""
    dev = mode_encoding.device
    weights = 2 ** torch.arange(9, dtype=torch.float32, device=dev)
    codes = (mode_encoding.float() * weights).sum(1)
    codes = codes[successes == 1]
    if codes.numel() == 0:
        return torch.zeros(0, device=dev), torch.zeros((), device=dev)
    _, counts = torch.unique(codes, return_counts=True)
    probs = counts.float() / counts.sum()
    ent = -(probs * (torch.log(probs) / torch.log(torch.tensor(24.0, device=dev)))).sum()
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
        prior_p = torch.full((keys.numel(),), 1.0 / keys.numel(), device=mode_encoding.device)
    else:
        keys = torch.tensor(list(prior.keys()), dtype=mode_encoding.dtype, device=mode_encoding.device)
        prior_p = torch.tensor([float(prior[k]) for k in prior.keys()], device=mode_encoding.device)
    n_mode, n_traj = keys.numel(), mode_encoding.shape[1]
    probs = torch.stack([((mode_encoding == k) & ok).sum(1).float() / n_traj for k in keys], 1)
    probs = probs / (probs.sum(1, keepdim=True) + 1e-12)
    probs = probs[probs.sum(1) != 0]
    if probs.shape[0] == 0:
        return probs, 0.0, 0.0
    base = torch.log(torch.tensor(float(max(n_mode, 2)), device=probs.device))
    entropy = -(probs * torch.log(probs + 1e-12) / base).sum(1).mean()
    log_ = (probs * torch.log(prior_p + 1e-12) / base).sum(1).mean()
    return probs, float(entropy), float(-entropy - log_)


def stacking_rows(info: torch.Tensor) -> torch.Tensor:
    """``Stacking_Sim.eval_agent`` result rows from the env's ``info`` words, on the device, without a per-env Python loop
    (``stacking_sim.py:122-141``): info = [success, mode string as base-4 digits (r 1, g 2, b 3; first arrival lowest),
    mean_distance, len(mode), status] -> [mode_3, mode_1, mode_2, success, success_1, success_2].

    MODE_1 = {r: 0, g: 1, b: 2}; MODE_2 = {rg: 0, rb: 1, gr: 2, gb: 3, br: 4, bg: 5}; MODE_3 = {rgb: 0, rbg: 1, grb: 2, gbr: 3,
    brg: 4, bgr: 5} — the third box of a 3-string is determined by the first two, so MODE_3 is indexed like MODE_2."""
    code, length = info[:, 1].long(), info[:, 3].long()
    d0, d1 = code % 4, (code // 4) % 4
    pair = (d0 - 1) * 2 + torch.where(d1 > d0, d1 - 2, d1 - 1)
    zero = torch.zeros_like(code)
    mode_1 = torch.where(length > 0, d0 - 1, zero)
    mode_2 = torch.where(length > 1, pair, zero)
    mode_3 = torch.where(length > 2, pair, zero)
    return torch.stack([mode_3, mode_1, mode_2, info[:, 0].long(), (length > 0).long(), (length > 1).long()], 1).to(info.dtype)
