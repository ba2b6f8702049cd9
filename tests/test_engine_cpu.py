"""End-to-end behaviour on CPU: local training parity with the reference Agent, learning / attack / defence shape
(qualitative README curves), PGD, partial participation, checkpoint/resume, logging."""
import copy
import json
import os
from types import SimpleNamespace

import pytest
import torch

from rlr_b200.engine import FLEngine
from rlr_b200.models import get_model
from rlr_b200.options import make_args
from rlr_b200.trainers import TorchTrainer
from rlr_b200.utils import load_checkpoint


def _engine(**kw):
    base = dict(data="fmnist", synthetic=1200, synthetic_val=300, num_agents=4, local_ep=1, bs=64, log_dir="", device="cpu")
    base.update(kw)
    return FLEngine(make_args(**base), verbose=False)


def test_local_step_matches_reference_agent_math():
    """One agent, no dropout randomness (eval of the same math): our flat fused step == torch SGD + clip_grad_norm_."""
    eng = _engine(model="cnn_cifar", data="cifar10", num_agents=1, synthetic=128, synthetic_val=64, bs=32)
    tr: TorchTrainer = eng.trainer
    agent = eng.agents[0]
    torch.manual_seed(5)
    w0 = eng.w_global.clone()
    out = torch.zeros_like(w0)
    torch.manual_seed(11)
    tr.train_agent(agent, eng.w_global, out, rnd=1)
    # replay with stock torch pieces on an independent copy
    net = get_model("cnn_cifar")
    net.w.copy_(w0)
    params = [p for p in net.parameters()]
    opt = torch.optim.SGD(params, lr=0.1, momentum=0.9)
    net.train()
    from rlr_b200.trainers import _agent_round_seed
    torch.manual_seed(_agent_round_seed(eng.args.seed, agent.id, 1))     # train_agent seeds the dropout stream per (agent, round)
    idx = agent.epoch_indices(eng.args.seed, 1, 0)
    for s in range(0, agent.n_data, 32):
        x, y = agent.dataset.batch(idx[s:s + 32])
        net.g.zero_()
        torch.nn.functional.cross_entropy(net(x), y).backward()
        torch.nn.utils.clip_grad_norm_(params, 10)
        opt.step()
    torch.testing.assert_close(out, net.w, atol=1e-5, rtol=1e-4)


def test_learns_and_rlr_defends_against_backdoor():
    """Qualitative shape of the README curves (performance.png / poison_acc.png): both runs learn the task; without
    the defence the backdoor is (intermittently, on this tiny synthetic set) learned, with RLR it stays near zero."""
    common = dict(num_agents=6, num_corrupt=1, poison_frac=1.0, local_ep=2, synthetic=1800, synthetic_val=400, seed=1)
    attacked = _engine(**common)
    defended = _engine(robustLR_threshold=3, **common)
    pa, pd, va, vd = [], [], [], []
    for r in range(1, 9):
        attacked.run_round(r)
        defended.run_round(r)
        if r >= 3:
            ea, ed = attacked.evaluate(r), defended.evaluate(r)
            pa.append(ea["poison_acc"]); pd.append(ed["poison_acc"]); va.append(ea["val_acc"]); vd.append(ed["val_acc"])
    assert max(va) > 0.9 and max(vd) > 0.9
    assert max(pa) > 0.5, "without the defence the backdoor gets in"
    assert sum(pd) / len(pd) < sum(pa) / len(pa) - 0.2, "RLR suppresses the backdoor"
    _, flipped = defended.round_result()
    assert 0 < flipped < defended.layout.n_vote


def test_pgd_keeps_update_inside_ball():
    eng = _engine(clip=0.05, num_agents=2)
    eng.run_round(1)
    for s in range(2):
        assert float((eng.fused.slots[s] - 0).norm()) > 0
    # slots still hold the local params of the round; global has moved by at most clip (mean of in-ball updates)
    eng2 = _engine(clip=0.05, num_agents=2)
    w0 = eng2.w_global.clone()
    eng2.run_round(1)
    assert float((eng2.w_global - w0)[: eng2.layout.n_vote].norm()) <= 0.05 + 1e-4


def test_partial_participation_and_sampling_is_seeded():
    eng = _engine(num_agents=10, agent_frac=0.3)
    assert eng.n_part == 3 and len(set(eng.sample_agents(1))) == 3
    assert eng.sample_agents(4) == eng.sample_agents(4) and eng.sample_agents(4) != eng.sample_agents(5)
    info = eng.run_round(1)
    assert len(info["chosen"]) == 3


def test_all_aggregators_run_with_noise():
    for aggr in ["avg", "comed", "sign"]:
        eng = _engine(aggr=aggr, server_lr=0.01, noise=0.001, clip=1.0, robustLR_threshold=2, num_agents=3)
        w0 = eng.w_global.clone()
        eng.run_round(1)
        assert torch.isfinite(eng.w_global).all() and not torch.equal(w0, eng.w_global)


def test_fit_logs_checkpoint_and_resume(tmp_path):
    ck = str(tmp_path / "ck.pt")
    eng = _engine(rounds=2, log_dir=str(tmp_path / "logs"), no_tensorboard=True, checkpoint=ck, snap=1)
    hist = eng.fit()
    eng.close()
    assert len(hist) == 2 and {"val_acc", "poison_acc", "train_loss", "frac_flipped", "ms_local_train", "ms_aggregate"} <= set(hist[-1])
    run_dirs = os.listdir(tmp_path / "logs")
    recs = [json.loads(l) for l in open(tmp_path / "logs" / run_dirs[0] / "metrics.jsonl")]
    assert [r["round"] for r in recs] == [1, 2]
    eng2 = _engine(rounds=3, resume=ck)
    assert eng2.start_round == 3
    torch.testing.assert_close(eng2.w_global, eng.w_global)
    assert len(eng2.fit()) == 1


def test_bn_model_round_on_cpu():
    eng = _engine(data="cifar10", model="vgg11", synthetic=128, synthetic_val=64, num_agents=2, bs=32)
    nv = eng.layout.n_vote
    stats0 = eng.w_global[nv:].clone()
    eng.run_round(1)
    assert not torch.equal(stats0, eng.w_global[nv:]), "BN running stats are aggregated (plain mean) too"
    assert eng.evaluate(1)["val_loss"] == eng.evaluate(1)["val_loss"]


def test_input_streaming_roundtrip():
    eng = _engine(num_agents=2)
    nbytes = eng.enable_input_streaming()
    assert nbytes == 1200 * (28 * 28 + 8)
    info = eng.run_round(1, stream_inputs=True)
    assert info["h2d_bytes"] == nbytes
    last = eng.agents[info["chosen"][-1]]                       # the staging buffer holds the shard uploaded last
    assert torch.equal(eng._stream_buf.data[: last.n_data], eng._stream_src[last.id][0])
    assert all(a.dataset is eng._stream_buf for a in eng.agents)


def test_fedemnist_shaped_float_data_path():
    """Fed-EMNIST stores float images in [0,1] (reference H5Dataset, src/utils.py:11-36): float gather path, `sub`-mode trojans
    and the non-IID-style agent count with partial participation (src/runner.sh:34-38 shape, scaled down)."""
    eng = _engine(data="fedemnist", synthetic=1500, synthetic_val=300, num_agents=30, agent_frac=0.2, num_corrupt=3, poison_frac=0.5,
                  pattern_type="square", robustLR_threshold=2, local_ep=2, bs=32)
    assert eng.train_dataset.data.dtype == torch.float32 and eng.n_part == 6
    assert any(len(a.poisoned_idxs) > 0 for a in eng.agents[:3]) and all(len(a.poisoned_idxs) == 0 for a in eng.agents[3:])
    for r in range(1, 4):
        eng.run_round(r)
    ev = eng.evaluate(3)
    assert ev["val_loss"] == ev["val_loss"] and 0.0 <= ev["poison_acc"] <= 1.0


def test_h5dataset_container_and_conversion():
    from rlr_b200.data import H5Dataset
    raw = {"c0": {"label": [1, 2, 3], "pixels": torch.rand(3, 28, 28).numpy()}, "c1": {"label": [4], "pixels": torch.rand(1, 28, 28).numpy()}}
    a, b = H5Dataset(raw, "c0"), H5Dataset(raw, "c1")
    assert len(a) == 3 and a[0][0].shape == (1, 28, 28) and set(a.classes().tolist()) == {1, 2, 3}
    ab = a + b
    assert len(ab) == 4 and ab.targets.tolist() == [1, 2, 3, 4]
    dd = ab.as_device_dataset()
    assert dd.data.shape == (4, 28, 28, 1) and dd.name == "fedemnist"


def test_participant_placement_balances_skewed_shards():
    """place_participants: same multiset, equal slot counts per rank, smaller maximum per-rank step sum than the sampled order for
    skewed shard sizes; untouched for equal shards and for a single process."""
    from types import SimpleNamespace
    from rlr_b200.engine import FLEngine
    import random
    rs = random.Random(0)
    sizes = [rs.choice([20, 40, 80, 160, 400, 900]) for _ in range(40)]
    fake = SimpleNamespace(ctx=SimpleNamespace(world=8), args=SimpleNamespace(bs=64, local_ep=2),
                           agents=[SimpleNamespace(n_data=n) for n in sizes])
    chosen = list(range(40)); rs.shuffle(chosen); chosen = chosen[:33]
    placed = FLEngine.place_participants(fake, chosen)
    assert sorted(placed) == sorted(chosen)
    steps = lambda a: 2 * ((sizes[a] + 63) // 64)
    load = lambda order: max(sum(steps(a) for a in order[r::8]) for r in range(8))
    assert load(placed) < load(chosen)
    ideal = sum(steps(a) for a in chosen) / 8
    assert load(placed) <= 1.35 * ideal
    assert FLEngine.place_participants(fake, placed) == FLEngine.place_participants(fake, chosen)      # deterministic, order-free
    fake.agents = [SimpleNamespace(n_data=100) for _ in sizes]
    assert FLEngine.place_participants(fake, chosen) == chosen                                           # equal shards: sampled order
    fake.ctx.world = 1
    fake.agents = [SimpleNamespace(n_data=n) for n in sizes]
    assert FLEngine.place_participants(fake, chosen) == chosen


def test_agents_in_flight_bookkeeping_matches_sequential():
    """--agents_in_flight: the agents of a rank are spread round-robin over several trainers (concurrent CUDA streams on a GPU); on
    CPU the schedule is still sequential, so the aggregated parameters and the round loss must equal the single-trainer run exactly
    for a dropout-free model."""
    res = {}
    for nf in (1, 3):
        eng = _engine(data="cifar10", model="resnet18", synthetic=160, synthetic_val=40, num_agents=5, local_ep=1, bs=16, agents_in_flight=nf,
                      robustLR_threshold=2)
        assert len(eng.trainers) == nf and eng.streams is None
        info = eng.run_round(1)
        loss, _ = eng.round_result()
        res[nf] = (eng.w_global.clone(), loss, info["steps"])
        eng.close()
    assert res[1][2] == res[3][2]
    torch.testing.assert_close(res[3][0], res[1][0], rtol=0, atol=0)
    assert abs(res[1][1] - res[3][1]) < 1e-4 * abs(res[1][1])          # per-trainer partial sums: float summation order only


@pytest.mark.parametrize("pgd", [0.0, 0.05])
def test_first_step_from_broadcast_buffer_equals_round_init_plus_step(pgd):
    """The hand-off fused with the first local step (ops.FlatSGD.step(w_in=...)): reading the parameters from the broadcast buffer
    with zero momentum must equal ``round_init`` (w <- w_global, m <- 0) followed by a normal step; coordinates behind ``n_pgd``
    (BatchNorm running statistics, already updated in ``w`` by the step's forward pass) keep their value."""
    from rlr_b200 import ops
    torch.manual_seed(3)
    n, n_vote = 4096, 3072
    w_global, g = torch.randn(n), torch.randn(n) * 3
    g[n_vote:] = 0                                              # buffers have no gradient
    stats = torch.randn(n - n_vote)                             # what the step's forward left in the trainer's running statistics
    # reference: separate round_init pass, then the step
    w_a, m_a = torch.zeros(n), torch.randn(n)
    ops.round_init(w_global, w_a, None, m_a)
    w_a[n_vote:] = stats
    ops.FlatSGD(n, "cpu", 0.1, 0.9, 10.0, pgd, n_pgd=n_vote).step(w_a, g, m_a, w0=w_global)
    # fused: stale w / m, parameters come from w_in
    w_b, m_b = torch.randn(n), torch.randn(n)
    w_b[n_vote:] = stats
    ops.FlatSGD(n, "cpu", 0.1, 0.9, 10.0, pgd, n_pgd=n_vote).step(w_b, g, m_b, w0=w_global, w_in=w_global)
    torch.testing.assert_close(w_b, w_a, rtol=1e-6, atol=1e-6)      # (w - lr*m vs w.add_(m, alpha=-lr): last-bit rounding only)
    torch.testing.assert_close(m_b, m_a, rtol=1e-6, atol=1e-6)


def test_agents_in_flight_rule_and_round_robin_on_cpu():
    """--agents_in_flight: 0 = auto (two trainers on the native CUDA path, one elsewhere); an explicit N builds N trainers that are used
    round-robin (on CPU without overlap) and changes nothing in the result."""
    a = _engine(rounds=1)
    assert len(a.trainers) == 1 and a.streams is None                                # CPU: auto = 1
    b = _engine(rounds=1, agents_in_flight=3)
    assert len(b.trainers) == min(3, b.fused.max_slots if hasattr(b.fused, "max_slots") else 3) and b.streams is None
    for r in (1, 2):
        a.run_round(r); b.run_round(r)
    torch.testing.assert_close(a.global_params(), b.global_params(), rtol=1e-5, atol=1e-6)
    a.close(); b.close()
