"""Observability / API-surface parity: TensorBoard tags (src/federated.py:81-91), stdout lines (:83-84,92), run-directory naming
(:27-30) and the reference's module-level names (utils.*, models.get_model, agent.Agent, aggregation.Aggregation, options.args_parser)."""
import glob
import os

import torch

from rlr_b200.engine import FLEngine
from rlr_b200.options import make_args

REF_TAGS = {"Validation/Loss", "Validation/Accuracy", "Poison/Base_Class_Accuracy", "Poison/Poison_Accuracy", "Poison/Poison_Loss",
            "Poison/Cumulative_Poison_Accuracy_Mean"}


def test_tensorboard_tags_and_stdout_lines(tmp_path, capsys):
    args = make_args(data="fmnist", synthetic=400, synthetic_val=100, num_agents=2, local_ep=1, bs=64, rounds=2, snap=1,
                     log_dir=str(tmp_path), device="cpu", num_corrupt=1, poison_frac=0.5, robustLR_threshold=2, diagnostics=True)
    eng = FLEngine(args, verbose=True)
    eng.fit()
    eng.close()
    out = capsys.readouterr().out
    assert "| Val_Loss/Val_Acc:" in out and "| Val_Per_Class_Acc:" in out and "| Poison Loss/Poison Acc:" in out
    assert "Training has finished!" in out and "Dataset: fmnist" in out
    run_dirs = os.listdir(tmp_path)
    assert len(run_dirs) == 1
    for frag in ("clip_val:", "noise_std:", "aggr:avg", "s_lr:1.0", "num_cor:1", "thrs_robustLR:2", "pttrn:plus"):
        assert frag in run_dirs[0]
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    acc = EventAccumulator(os.path.join(tmp_path, run_dirs[0]))
    acc.Reload()
    tags = set(acc.Tags()["scalars"])
    assert REF_TAGS <= tags
    assert {"Norms/Avg_Honest_L2", "Norms/Avg_Corrupt_L2"} <= tags, "the reference's disabled diagnostics are available behind --diagnostics"
    assert [e.step for e in acc.Scalars("Validation/Accuracy")] == [1, 2]
    assert glob.glob(os.path.join(tmp_path, run_dirs[0], "metrics.jsonl"))


def test_reference_module_names_resolve():
    from rlr_b200 import agent, aggregation, models, options, utils
    for name in ("get_datasets", "distribute_data", "poison_dataset", "add_pattern_bd", "get_loss_n_accuracy", "DatasetSplit", "H5Dataset",
                 "print_exp_details"):
        assert callable(getattr(utils, name)), name
    assert callable(models.get_model) and callable(options.args_parser)
    assert hasattr(agent.Agent, "local_train") and hasattr(aggregation.Aggregation, "aggregate_updates")
    for m in ("compute_robustLR", "agg_avg", "agg_comed", "agg_sign", "clip_updates", "plot_norms", "comp_diag_fisher", "plot_sign_agreement"):
        assert hasattr(aggregation.Aggregation, m), m
    net = models.get_model("fmnist")
    assert net(torch.zeros(1, 1, 28, 28)).shape == (1, 10)
