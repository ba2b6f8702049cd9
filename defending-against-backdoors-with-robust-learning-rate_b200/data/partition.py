"""Agent data partitioner (reference ``distribute_data``, src/utils.py:58-92).

Semantics reproduced exactly (differentially tested against the reference in tests/test_partition.py):
labels are sorted, every class's index list is cut into ``slice_size`` strided chunks
(``v[i::slice_size]``), and agent ``u`` takes the first remaining chunk of each class in class order until it
holds ``class_per_agent`` chunks.  ``class_per_agent == n_classes`` gives the IID split the reference always
uses; smaller values give a label-skewed (non-IID) split.  ``num_agents == 1`` -> the whole dataset.
"""
from __future__ import annotations

from collections import defaultdict


def distribute_data(dataset, args=None, n_classes: int = 10, class_per_agent: int = 10, num_agents: int | None = None):
    """Return ``{agent_id: list[int]}`` of sample indices."""
    K = num_agents if num_agents is not None else args.num_agents
    n = len(dataset)
    if K == 1:
        return {0: range(n)}
    # same call as the reference (src/utils.py:66) so tie order inside a class is identical
    order = dataset.targets.detach().cpu().sort()
    labels, indices = order.values.tolist(), order.indices.tolist()
    by_class = defaultdict(list)
    for lab, idx in zip(labels, indices):
        by_class[lab].append(idx)

    shard_size = n // (K * class_per_agent)
    if shard_size == 0:
        raise ValueError(f"dataset of {n} samples is too small for {K} agents x {class_per_agent} classes")
    slice_size = (n // n_classes) // shard_size
    chunks = {c: [v[i::slice_size] for i in range(slice_size)] for c, v in by_class.items()}

    users = defaultdict(list)
    for u in range(K):
        taken = 0
        for c in range(n_classes):
            if taken == class_per_agent:
                break
            if len(chunks.get(c, ())) > 0:
                users[u] += chunks[c].pop(0)
                taken += 1
    return users
