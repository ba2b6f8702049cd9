"""Aggregation server (reference ``Aggregation``, src/aggregation.py:8-190).

``aggregate_updates`` = Robust-LR sign vote + {FedAvg | coordinate median | sign majority} + optional noise + server
step, executed by ONE fused kernel over the flat parameter vector (``ops.fused_aggregate`` in-process, or
``parallel.FusedAggregator`` across GPUs) instead of the reference's ~30 elementwise fp64 passes.
The reference's dead / disabled pieces are available behind flags: ``clip_updates`` (``--server_clip``) and the
diagnostics ``plot_norms`` / ``comp_diag_fisher`` / ``plot_sign_agreement`` (``--diagnostics``), the latter with the
reference's latent bugs fixed (model built on the right device; Fisher uses log-probabilities -- SURVEY.md quirk 7).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import ops
from .models.graph import GraphNet


class Aggregation:
    def __init__(self, agent_data_sizes, n_params, poisoned_val, args, writer=None, layout=None, fused=None):
        self.agent_data_sizes = agent_data_sizes
        self.args = args
        self.writer = writer
        self.server_lr = args.server_lr
        self.n_params = n_params
        self.poisoned_val = poisoned_val
        self.layout = layout
        self.fused = fused            # parallel.FusedAggregator or None (pure in-process use)
        self.cum_net_mov = 0.0
        self.last_flipped = 0

    # ---- the server step ------------------------------------------------------------------------------------
    def aggregate_updates(self, w_global, agent_params, cur_round, n_vote=None):
        """In-process form: ``agent_params`` = {agent_id: flat local parameters}.  Updates ``w_global`` in place."""
        ids = list(agent_params.keys())
        ws = [agent_params[i] for i in ids]
        weights = [float(self.agent_data_sizes[i]) for i in ids]
        nv = n_vote if n_vote is not None else (self.layout.n_vote if self.layout else None)
        scales = self._clip_scales(ops.update_norms(w_global, ws, nv)) if self._server_clip else None
        prev = w_global.clone() if self.args.diagnostics else None
        flipped = torch.zeros(1, dtype=torch.int64, device=w_global.device)
        ops.fused_aggregate(w_global, ws, weights, self.args.aggr, self.args.robustLR_threshold, self.server_lr,
                            self.args.noise * self.args.clip, self.args.seed, cur_round,
                            n_vote if n_vote is not None else (self.layout.n_vote if self.layout else None),
                            scales, out=w_global, flipped=flipped)
        self.last_flipped = flipped
        if self.args.diagnostics:
            self.plot_norms(dict(zip(ids, ops.update_norms(prev, ws, nv).tolist())), cur_round)
            self.plot_sign_agreement(prev, w_global, ws, ids, cur_round)
        return

    def aggregate_slots(self, participants, cur_round):
        """Engine form: participant j's parameters live in ``fused.slot_owner(j)``; updates every rank's global."""
        weights = [float(self.agent_data_sizes[i]) for i in participants]
        scales = None
        norms = None
        diag = bool(self.args.diagnostics)
        if self._server_clip or diag:
            norms = self.fused.update_norms(len(participants))
        if self._server_clip:
            scales = self._clip_scales(norms)
        if diag:   # the sign-agreement analysis needs the pre-step global parameters and every participant's parameters
            prev = self.fused.w_global.clone()
            ws = [w.clone() for w in self.fused.gather_participants(len(participants))]
        self.fused.aggregate(weights, self.args.aggr, self.args.robustLR_threshold, self.server_lr,
                             self.args.noise * self.args.clip, self.args.seed, cur_round, scales)
        self.last_flipped = self.fused.flipped
        if diag:
            self.plot_norms(dict(zip(participants, norms.tolist())), cur_round)
            self.plot_sign_agreement(prev, self.fused.w_global, ws, participants, cur_round)

    @property
    def _server_clip(self):
        return bool(getattr(self.args, "server_clip", False)) and self.args.clip > 0

    def _clip_scales(self, norms):
        """reference ``clip_updates`` (src/aggregation.py:77-81): update /= max(1, ||update||/clip)."""
        return (1.0 / torch.clamp(norms / self.args.clip, min=1.0)).float()

    # ---- reference-named helpers (thin wrappers over the oracle; kept for API parity and tests) -----------------
    def compute_robustLR(self, agent_updates_dict):
        """±server_lr per coordinate from the sign vote (src/aggregation.py:48-54)."""
        s = sum(torch.sign(u) for u in agent_updates_dict.values()).abs()
        return torch.where(s >= self.args.robustLR_threshold, self.server_lr, -self.server_lr).to(s.dtype)

    def agg_avg(self, agent_updates_dict):
        tot = sum(self.agent_data_sizes[i] for i in agent_updates_dict)
        return sum(self.agent_data_sizes[i] * u for i, u in agent_updates_dict.items()) / tot

    def agg_comed(self, agent_updates_dict):
        return torch.median(torch.stack(list(agent_updates_dict.values()), dim=1), dim=1).values

    def agg_sign(self, agent_updates_dict):
        return torch.sign(sum(torch.sign(u) for u in agent_updates_dict.values()))

    def clip_updates(self, agent_updates_dict):
        for u in agent_updates_dict.values():
            u.div_(max(1.0, float(torch.norm(u, p=2)) / self.args.clip))

    # ---- diagnostics ---------------------------------------------------------------------------------------
    def plot_norms(self, norms_by_agent, cur_round, norm=2):
        """Average update norm of honest vs corrupt agents (src/aggregation.py:83-100)."""
        honest = [v for k, v in norms_by_agent.items() if k >= self.args.num_corrupt]
        corrupt = [v for k, v in norms_by_agent.items() if k < self.args.num_corrupt]
        out = {}
        if honest:
            out[f"Norms/Avg_Honest_L{norm}"] = sum(honest) / len(honest)
        if corrupt:
            out[f"Norms/Avg_Corrupt_L{norm}"] = sum(corrupt) / len(corrupt)
        for k, v in out.items():
            if self.writer is not None:
                self.writer.add_scalar(k, v, cur_round)
        self.last_norms = out
        return out

    def comp_diag_fisher(self, model_params, dataset, adv=True, bs=256):
        """Diagonal Fisher information of the log-likelihood of the (adversarial or base-class) label on the poisoned
        validation set (src/aggregation.py:102-129, with quirk 7 fixed)."""
        dev = model_params.device
        w = model_params.clone()
        g = torch.zeros_like(w)
        net = GraphNet(self.layout, w, g)
        net.eval()
        fisher = torch.zeros_like(w)
        n = len(dataset)
        for start in range(0, n, bs):
            idx = torch.arange(start, min(n, start + bs), device=dev)
            x, y = dataset.batch(idx)
            if not adv:
                y = torch.full_like(y, self.args.base_class)
            g.zero_()
            logp = F.log_softmax(net(x), dim=1)
            logp.gather(1, y[:, None]).sum().backward()
            fisher += g ** 2 / n
        return fisher[: self.layout.n_vote].detach()

    def plot_sign_agreement(self, cur_global_params, new_global_params, agent_params, ids, cur_round):
        """Which of the most backdoor-relevant coordinates (top-``top_frac`` Fisher) had their LR kept vs flipped, for
        adversarial vs honest objectives; logs the 7 ``Sign/*`` scalars (src/aggregation.py:132-190)."""
        if self.layout is None or self.poisoned_val is None or len(self.poisoned_val) == 0:
            return {}
        nv = self.layout.n_vote
        update = (new_global_params - cur_global_params)[:nv]
        signs = sum(torch.sign(w[:nv] - cur_global_params[:nv]) for w in agent_params).abs()
        theta = self.args.robustLR_threshold
        lr = torch.where(signs >= theta, 1.0, -1.0) if theta > 0 else torch.ones_like(signs)
        fa = self.comp_diag_fisher(cur_global_params, self.poisoned_val, adv=True)
        fh = self.comp_diag_fisher(cur_global_params, self.poisoned_val, adv=False)
        k = int(self.args.top_frac)
        adv_top = fa.topk(k).indices.cpu().numpy()
        hon_top = fh.topk(k).indices.cpu().numpy()
        min_idxs = (lr < 0).nonzero().flatten().cpu().numpy()
        max_idxs = (lr > 0).nonzero().flatten().cpu().numpy()
        max_adv, max_hon = np.intersect1d(adv_top, max_idxs), np.intersect1d(hon_top, max_idxs)
        min_adv, min_hon = np.intersect1d(adv_top, min_idxs), np.intersect1d(hon_top, min_idxs)
        l2 = lambda ix: float(torch.norm(update[torch.as_tensor(ix, dtype=torch.int64, device=update.device)])) if len(ix) else 0.0
        v = {
            "Sign/Hon_Maxim_L2": l2(np.setdiff1d(max_hon, max_adv)), "Sign/Adv_Maxim_L2": l2(np.setdiff1d(max_adv, max_hon)),
            "Sign/Adv_Minim_L2": l2(np.setdiff1d(min_adv, min_hon)), "Sign/Hon_Minim_L2": l2(np.setdiff1d(min_hon, min_adv)),
        }
        v["Sign/Adv_Net_L2"] = v["Sign/Adv_Maxim_L2"] - v["Sign/Adv_Minim_L2"]
        v["Sign/Hon_Net_L2"] = v["Sign/Hon_Maxim_L2"] - v["Sign/Hon_Minim_L2"]
        self.cum_net_mov += v["Sign/Hon_Net_L2"] - v["Sign/Adv_Net_L2"]
        v["Sign/Model_Net_L2_Cumulative"] = self.cum_net_mov
        for key, val in v.items():
            if self.writer is not None:
                self.writer.add_scalar(key, val, cur_round)
        self.last_sign_stats = v
        return v
