"""Batched use of the reference's *unchanged* agents inside the rollout loop.

The reference calls ``agent.predict(np.ndarray[obs_dim]) -> np.ndarray[1, act_dim]`` once per env step per process
(``agents/base_agent.py:110-122``).  For N lock-stepped envs ``predict_batch`` dispatches on the agent's CLASS:

  * an agent that brings its own ``predict_batch(obs[N, obs_dim])`` is used as is;
  * ``BC_Agent`` (``agents/bc_agent.py:241-271``): scale -> model -> clamp -> inverse scale on an [N, 1, obs] tensor;
  * ``DiffusionAgent`` (``agents/ddpm_agent.py:214-274``): scale, window of past observations (``obs_context`` now holds
    [N, obs] rows), EMA parameter swap, model = full reverse-diffusion sampler, restore, inverse scale;
  * anything else must be stateless to be looped per env; an agent with per-episode state (``obs_context``, ``action_counter``,
    ``curr_action_seq`` ... — ACT, BeT, BESO, GPT-BC, CVAE) and no batched path raises ``NotImplementedError`` instead of
    silently sharing ONE episode state across N envs.

Duck-typed dispatch (``model`` + ``scaler`` + ...) is deliberately not used: it matches ACT / BeT / BESO / the CVAE agent as
well, whose ``predict`` does more than ``model(x)`` (action chunking, latent + offset decoding, prior sampling).
A stand-in that follows one of the two supported laws can declare it with ``batched_kind = "bc" | "ddpm"``.
"""
from __future__ import annotations

import numpy as np
import torch

_STATEFUL_ATTRS = ("obs_context", "action_counter", "curr_action_seq", "action_context", "que_actions", "pre_obs")


def _kind(agent) -> str | None:
    k = getattr(agent, "batched_kind", None)
    if k in ("bc", "ddpm"):
        return k
    names = {c.__name__ for c in type(agent).__mro__}
    if "DiffusionAgent" in names and not getattr(agent, "diffusion_kde", False):
        return "ddpm"
    if "BC_Agent" in names:
        return "bc"
    return None


def graph_safe(agent) -> bool:
    """True when ``predict_batch(agent, .)`` is pure tensor code with no Python-side state between calls, i.e. the rollout body can
    be captured in a CUDA graph: agents that say so (``graph_safe = True``: the synthetic policies), the BC path, and the
    diffusion path without an observation window (the deque of past observations is Python state a graph replay would skip)."""
    if getattr(agent, "graph_safe", None) is not None:
        return bool(agent.graph_safe)
    if hasattr(agent, "predict_batch"):
        return False
    kind = _kind(agent)
    dev = str(getattr(agent, "device", "cuda"))
    return dev.startswith("cuda") and (kind == "bc" or (kind == "ddpm" and getattr(agent, "window_size", 1) <= 1))


@torch.no_grad()
def predict_batch(agent, obs: torch.Tensor) -> torch.Tensor:
    """obs: [N, obs_dim] float tensor (any device). Returns [N, act_dim] float32 on obs.device."""
    if hasattr(agent, "predict_batch"):
        return torch.as_tensor(agent.predict_batch(obs), dtype=torch.float32, device=obs.device)
    kind = _kind(agent)
    if kind == "bc":
        agent.model.eval()
        dev = getattr(agent, "device", obs.device)
        x = obs.to(dev).float().unsqueeze(1)                 # [N, 1, obs] like predict()'s unsqueeze(0).unsqueeze(0) per sample
        x = agent.scaler.scale_input(x)
        out = agent.model(x)
        out = out.clamp_(agent.min_action, agent.max_action)
        out = agent.scaler.inverse_scale_output(out)
        return out[:, 0].to(obs.device, torch.float32)
    if kind == "ddpm":
        dev = getattr(agent, "device", obs.device)
        state = agent.scaler.scale_input(obs.to(dev).float())
        if getattr(agent, "window_size", 1) > 1:
            agent.obs_context.append(state)                  # deque(maxlen=window_size), cleared by agent.reset(): rows are envs now
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
    stateful = [a for a in _STATEFUL_ATTRS if hasattr(agent, a)]
    if stateful:
        raise NotImplementedError(
            f"{type(agent).__name__} keeps per-episode state ({', '.join(stateful)}) and has no batched path: looping its predict() over "
            f"N envs would share one episode state between them.  Give it a predict_batch(obs[N, obs_dim]) method.")
    acts = [np.asarray(agent.predict(o))[0] for o in obs.detach().cpu().numpy()]
    return torch.as_tensor(np.stack(acts), dtype=torch.float32, device=obs.device)
