"""Fused aggregate + server-step + broadcast across GPUs (the product path of SURVEY.md 5.8).

Per-rank symmetric slab layout (byte offsets identical on every rank):

    [ flags: uint32[3*world] (+pad to 4 KiB) | w_global fp32[n] | w_global bf16[n] | slot_0 fp32[n] | slot_1 ... ]

Flag words of a rank: [0, world) barrier-in arrivals, [world, 2 world) barrier-out arrivals, [2 world, 3 world) broadcast-ready
words ("slice r of the new global parameters has landed here", written by rank r).

``w_global`` is what every local trainer reads at the start of a round; ``slot_j`` receives the parameters of the
j-th agent this rank trained.  One launch of ``fused_aggregate_kernel`` per rank then (i) waits until every rank has
signalled "slots ready", (ii) reads the owned coordinate slice of every participant's slot straight from the owning
GPU's HBM over NVLink, (iii) computes sign vote / aggregator / RLR flip / server step, (iv) stores the new global
slice into EVERY rank's ``w_global`` (+bf16 shadow) with NVLS multicast stores (or per-peer P2P stores), and (v)
signals/awaits "slice landed".  No NCCL call, no host synchronisation, no materialised update vectors.

Reference counterpart: the Python dict ``agent_updates_dict[agent_id] = update`` (src/federated.py:67-70), the ~30 elementwise
fp64 passes of ``Aggregation.aggregate_updates`` (src/aggregation.py:19-75) and the per-agent
``vector_to_parameters(copy.deepcopy(rnd_global_params), ...)`` "broadcast" (src/federated.py:72).

Hand-off fused with the next round's first GEMM (``enable_handoff``): the kernel then has NO barrier-out.  Rank r
publishes its slice in every peer's ready word and exits; the consumer of the broadcast is the first local step of the next round
(``models.native.NativeTrainer``): the producer warp of the stem convolution's tcgen05 GEMM acquires the ready word(s) of the slice(s)
that hold its filter -- it reads that filter straight out of the multicast bf16 shadow -- and a one-warp ``acquire_slices`` kernel
queued right behind it waits for the remaining slices, so the first-layer GEMM overlaps the rest of the broadcast.  The separate
``round_init`` pass (w <- w_global, bf16 shadow, momentum <- 0) does not exist on that path: the first optimizer step reads
``w_global`` directly with zero momentum.  Host-side readers of ``w_global`` (evaluation, checkpoints, tests) call ``acquire()`` first.

With ``backend in {nccl, gloo}`` (the baseline transport) the slots are all-gathered and the same kernel runs on the
gathered copies locally; on CPU it runs the fp64 oracle.
"""
from __future__ import annotations

import torch

from .. import ops
from .symm import SymmetricBuffer

FLAG_BYTES = 4096


class FusedAggregator:
    def __init__(self, ctx, n_total: int, n_vote: int, max_slots: int, backend: str = "auto", with_bf16: bool = True,
                 transport: str = "auto"):
        self.ctx = ctx
        # nccl / gloo back-ends: "gather" all-gathers every participant's parameters and runs the kernel on the copies (any
        # aggregator); "reduce" all-reduces per-coordinate partial sums (vote, weighted update sum) -- O(N) instead of O(K N)
        # traffic per rank, avg / sign / RLR only (coordinate median falls back to gather).  auto = reduce across hosts.
        self.transport = transport if transport in ("gather", "reduce") else ("gather" if getattr(ctx, "single_node", True) else "reduce")
        self.n, self.n_vote, self.max_slots = int(n_total), int(n_vote), int(max_slots)
        dev = ctx.device
        if backend == "auto":
            backend = "fused" if (dev.type == "cuda") else ("gloo" if ctx.is_dist else "local")
            if backend == "fused" and ctx.is_dist and not getattr(ctx, "single_node", True):
                backend = "nccl"       # ranks on several hosts: no common peer-memory domain -> all-gather transport + local kernel
                if ctx.is_main:
                    print("[parallel] ranks span several hosts: aggregation uses the NCCL all-gather transport")
        if backend == "fused" and ctx.is_dist and not getattr(ctx, "single_node", True):
            raise ValueError("backend=fused needs all ranks on one host (peer-mapped memory); use --backend nccl across hosts")
        if backend == "fused" and dev.type != "cuda":
            raise ValueError("backend=fused needs CUDA devices")
        self.backend = backend
        self.with_bf16 = with_bf16 and dev.type == "cuda"
        n = self.n
        self.off_flags = 0
        self.off_wg = FLAG_BYTES
        self.off_wb = self.off_wg + 4 * n
        self.off_slots = self.off_wb + 2 * n
        nbytes = self.off_slots + 4 * n * self.max_slots
        use_symm = backend == "fused" and ctx.is_dist
        self.buf = SymmetricBuffer(ctx, nbytes) if use_symm else SymmetricBuffer(_Solo(ctx), nbytes)
        self.w_global = self.buf.tensor(self.off_wg, n, torch.float32)
        self.w_bf16 = self.buf.tensor(self.off_wb, n, torch.bfloat16) if self.with_bf16 else None
        self.slots = [self.buf.tensor(self.off_slots + 4 * n * j, n, torch.float32) for j in range(self.max_slots)]
        self.flipped = torch.zeros(1, dtype=torch.int64, device=dev)
        self.flipped_is_partial = False   # True after a launch in which every rank counted only its own coordinate slice
        self.epoch = 0
        self._tables = {}
        # hand-off state: the epoch consumers of the broadcast wait for (device word read by captured kernels), ready-word address
        self.handoff = False
        self.epoch_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.ready_ptr = 0
        self.n_slices = 0
        self.per = n
        if use_symm:
            world = ctx.world
            self.local_sync = torch.zeros(2, dtype=torch.int32, device=dev)
            self.flag_ptrs = ops.PtrTable([self.buf.peer_ptr(r, self.off_flags) for r in range(world)], dev)
            mc = self.buf.mc_ptr(self.off_wg)
            self.use_multimem = bool(mc)
            if self.use_multimem:
                self.out_ptrs = ops.PtrTable([mc], dev)
                self.out_bf16_ptrs = ops.PtrTable([self.buf.mc_ptr(self.off_wb)], dev) if self.with_bf16 else None
            else:
                self.out_ptrs = ops.PtrTable([self.buf.peer_ptr(r, self.off_wg) for r in range(world)], dev)
                self.out_bf16_ptrs = (ops.PtrTable([self.buf.peer_ptr(r, self.off_wb) for r in range(world)], dev)
                                      if self.with_bf16 else None)
            # coordinate slices: multiples of 4, cover [0, n)
            per = (n // 4 + world - 1) // world * 4
            self.per = per
            self.begin = min(n, ctx.rank * per)
            self.end = min(n, self.begin + per)

    # ---- hand-off fused with the next round's first GEMM -------------------------------------------------------------------
    def enable_handoff(self):
        """Replace the kernel's barrier-out by per-slice ready words that the consumers acquire (see the module docstring).  With
        one process (no peers) there is nothing to wait for: consumers get ``ready_ptr = 0`` and only the round_init-free first step
        remains.  Returns True."""
        self.handoff = True
        if self.backend == "fused" and self.ctx.is_dist:
            world = self.ctx.world
            self.ready_ptr = self.buf.peer_ptr(self.ctx.rank, self.off_flags) + 4 * 2 * world
            self.n_slices = world
        return True

    def slices_of(self, lo: int, hi: int):
        """Indices (first, last) of the broadcast slices that hold coordinates [lo, hi)."""
        if not self.n_slices:
            return 0, 0
        return min(self.n_slices - 1, lo // self.per), min(self.n_slices - 1, max(lo, hi - 1) // self.per)

    def acquire(self):
        """Make every slice of the current global parameters visible to work queued afterwards on the current stream (needed
        before reading ``w_global`` outside the trainers when the hand-off is fused; a no-op otherwise)."""
        if self.handoff and self.ready_ptr:
            ops.ext().acquire_slices(self.ready_ptr, 0, self.n_slices - 1, self.epoch_dev, None, None)

    # ---------------------------------------------------------------------------------------------------------
    def slot_owner(self, j: int):
        """(rank, local slot) of the j-th participant of a round."""
        return j % self.ctx.world, j // self.ctx.world

    def _agent_table(self, n_part: int):
        if n_part not in self._tables:
            ptrs = []
            for j in range(n_part):
                r, s = self.slot_owner(j)
                ptrs.append(self.buf.peer_ptr(r, self.off_slots + 4 * self.n * s))
            self._tables[n_part] = ops.PtrTable(ptrs, self.ctx.device)
        return self._tables[n_part]

    def aggregate(self, weights, mode, theta, server_lr, noise_std=0.0, seed=0, rnd=0, scales=None):
        """Aggregate the first ``len(weights)`` participants (participant j lives in ``slot_owner(j)``) and update
        ``w_global`` on every rank.  Returns nothing; ``self.flipped`` accumulates the flipped-coordinate count."""
        n_part = len(weights)
        ctx, dev = self.ctx, self.ctx.device
        self.flipped.zero_()
        self.flipped_is_partial = False
        if self.backend == "fused" and ctx.is_dist and n_part <= ops.MAX_FUSED_AGENTS:
            self.flipped_is_partial = True
            self.epoch += 1
            wt = torch.as_tensor(weights, dtype=torch.float64).to(dev)
            sc = torch.as_tensor(scales, dtype=torch.float32).to(dev) if scales is not None else None
            ops.ext().fused_aggregate(
                self._agent_table(n_part).tensor, wt, sc, float(sum(float(x) for x in weights)), self.w_global.data_ptr(),
                self.out_ptrs.tensor, self.out_bf16_ptrs.tensor if self.out_bf16_ptrs else None, self.use_multimem,
                self.begin, self.end, self.n_vote, ops.MODE_IDS[mode], int(theta), float(server_lr), float(noise_std),
                int(seed), int(rnd), self.flipped, self.flag_ptrs.tensor, self.local_sync, ctx.rank, ctx.world, self.epoch,
                bool(self.handoff))
            if self.handoff:
                self.epoch_dev.fill_(self.epoch)      # what the next round's consumers wait for (stream-ordered before their graphs)
            return
        # ---- baseline transports / single process ------------------------------------------------------------------
        if ctx.is_dist and self.transport == "reduce" and mode in ("avg", "sign"):
            self._aggregate_reduce(weights, mode, theta, server_lr, noise_std, seed, rnd, scales)
            return
        # gather participant params, run the kernel locally
        if ctx.is_dist:
            mine = torch.stack(self.slots, 0)                       # [max_slots, n]
            allp = ctx.all_gather(mine)                             # [world, max_slots, n]
            agents = [allp[j % ctx.world, j // ctx.world] for j in range(n_part)]
        else:
            agents = [self.slots[j] for j in range(n_part)]
        ops.fused_aggregate(self.w_global, agents, weights, mode, theta, server_lr, noise_std, seed, rnd, self.n_vote,
                            scales, out=self.w_global, out_bf16=self.w_bf16, flipped=self.flipped)

    def _aggregate_reduce(self, weights, mode, theta, server_lr, noise_std, seed, rnd, scales):
        """All-reduce transport: every rank folds its local participants into (vote, weighted sum), two all_reduce calls make them
        global, and every rank finishes the identical server step locally (ops.aggregate_from_partials).  The noise vector is
        drawn from a torch generator seeded like the CPU oracle on every rank (same stream everywhere; it is NOT the Philox
        stream of the fused kernel)."""
        ctx, n_part = self.ctx, len(weights)
        mine = [j for j in range(n_part) if self.slot_owner(j)[0] == ctx.rank]
        vote, wsum = ops.aggregate_partials(self.w_global, [self.slots[self.slot_owner(j)[1]] for j in mine], [weights[j] for j in mine],
                                            self.n_vote, [scales[j] for j in mine] if scales is not None else None)
        ctx.all_reduce_sum(vote)
        ctx.all_reduce_sum(wsum)
        noise = None
        if noise_std > 0:
            gen = torch.Generator().manual_seed(int(seed) * 1000003 + int(rnd))
            noise = (torch.randn(self.n, generator=gen, dtype=torch.float64) * noise_std).to(self.w_global.device)
            noise[self.n_vote:] = 0
        new, nflip = ops.aggregate_from_partials(self.w_global, vote, wsum, sum(float(x) for x in weights), mode, theta, server_lr,
                                                 noise, self.n_vote)
        self.w_global.copy_(new)
        if self.w_bf16 is not None:
            self.w_bf16.copy_(new.to(torch.bfloat16))
        self.flipped += nflip

    def gather_participants(self, n_part: int):
        """Every participant's flat parameter vector on THIS rank (list of ``n_part`` tensors; remote slots are copied through an
        all_gather).  Off the hot path: used by the ``--diagnostics`` analyses and by bench.py's post-run aggregation check."""
        ctx = self.ctx
        if not ctx.is_dist:
            return [self.slots[j] for j in range(n_part)]
        allp = ctx.all_gather(torch.stack(self.slots, 0))            # [world, max_slots, n]
        return [allp[j % ctx.world, j // ctx.world] for j in range(n_part)]

    def update_norms(self, n_part: int):
        """||w_j - w_global|| for every participant (float64 [n_part]), computed where the slot lives."""
        ctx = self.ctx
        mine = [j for j in range(n_part) if self.slot_owner(j)[0] == ctx.rank]
        local = torch.zeros(n_part, dtype=torch.float64, device=ctx.device)
        if mine:
            norms = ops.update_norms(self.w_global, [self.slots[self.slot_owner(j)[1]] for j in mine], self.n_vote)
            local[torch.as_tensor(mine, device=ctx.device)] = norms ** 2
        ctx.all_reduce_sum(local)
        return local.sqrt()

    def close(self):
        self.buf.close()


class _Solo:
    """Context stand-in that makes SymmetricBuffer allocate plain local memory."""

    def __init__(self, ctx):
        self.device, self.world, self.rank, self.is_dist, self.is_main = ctx.device, 1, 0, False, True
