"""Batched use of the reference's *unchanged* agents inside the rollout loop.

The reference calls ``agent.predict(np.ndarray[obs_dim]) -> np.ndarray[1, act_dim]`` once per env step per process
(``agents/base_agent.py:110-122``).  For N lock-stepped envs we either
  * use a batched path that mirrors ``BC_Agent.predict`` (``agents/bc_agent.py:241-271``: scale -> model -> clamp ->
    inverse scale) on a [N, 1, obs] tensor, reading only public attributes (``model``, ``scaler``, ``min_action``,
    ``max_action``, ``device``), or
  * fall back to looping ``agent.predict`` per env — only valid for agents without per-episode state.
"""
from __future__ import annotations

import numpy as np
import torch


def _is_bc_like(agent) -> bool:
    return all(hasattr(agent, a) for a in ("model", "scaler", "min_action", "max_action")) and not hasattr(agent, "obs_context")


def _is_ddpm_like(agent) -> bool:
    return all(hasattr(agent, a) for a in ("model", "scaler", "obs_context")) and not getattr(agent, "diffusion_kde", False)


@torch.no_grad()
def predict_batch(agent, obs: torch.Tensor) -> torch.Tensor:
    """obs: [N, obs_dim] float tensor (any device). Returns [N, act_dim] float32 on obs.device."""
    if hasattr(agent, "predict_batch"):
        return torch.as_tensor(agent.predict_batch(obs), dtype=torch.float32, device=obs.device)
    if _is_bc_like(agent):
        agent.model.eval()
        dev = getattr(agent, "device", obs.device)
        x = obs.to(dev).float().unsqueeze(1)                 # [N, 1, obs] like predict()'s unsqueeze(0).unsqueeze(0) per sample
        x = agent.scaler.scale_input(x)
        out = agent.model(x)
        out = out.clamp_(agent.min_action, agent.max_action)
        out = agent.scaler.inverse_scale_output(out)
        return out[:, 0].to(obs.device, torch.float32)
    if _is_ddpm_like(agent):
        # DiffusionAgent.predict (agents/ddpm_agent.py:214-274) on N rows at once: scale, (window of past observations),
        # EMA parameter swap, model = full reverse-diffusion sampler, restore, inverse scale.  agent.reset() clears the window.
        dev = getattr(agent, "device", obs.device)
        state = agent.scaler.scale_input(obs.to(dev).float())
        if getattr(agent, "window_size", 1) > 1:
            agent.obs_context.append(state)
            inp = torch.stack(tuple(agent.obs_context), dim=1)
        else:
            inp = state
        ema = getattr(agent, "use_ema", False)
        if ema:
            agent.ema_helper.store(agent.model.parameters())
            agent.ema_helper.copy_to(agent.model.parameters())
        agent.model.eval()
        pred = agent.model(inp, None)
        if pred.dim() == 3:
            pred = pred[:, -1, :]
        if ema:
            agent.ema_helper.restore(agent.model.parameters())
        return agent.scaler.inverse_scale_output(pred).to(obs.device, torch.float32)
    acts = [np.asarray(agent.predict(o))[0] for o in obs.detach().cpu().numpy()]
    return torch.as_tensor(np.stack(acts), dtype=torch.float32, device=obs.device)
