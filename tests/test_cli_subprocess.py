"""The command-line entry points run as the reference's would (``python federated.py --flags``), on CPU."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=300):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, *args], cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


def test_federated_py_reference_style_command_line(tmp_path):
    r = _run(["federated.py", "--data=fmnist", "--local_ep=1", "--bs=64", "--num_agents=3", "--rounds=2", "--num_corrupt=1",
              "--poison_frac=0.5", "--robustLR_threshold=2", "--synthetic=300", "--synthetic_val=60", f"--log_dir={tmp_path}",
              "--no_tensorboard", "--device=cpu"])
    assert r.returncode == 0, r.stderr[-2000:]
    assert "| Val_Loss/Val_Acc:" in r.stdout and "| Poison Loss/Poison Acc:" in r.stdout and "Training has finished!" in r.stdout
    assert "Aggregation Function: avg" in r.stdout and "RobustLR_threshold: 2" in r.stdout


def test_module_entry_point_and_help():
    r = _run(["-m", "rlr_b200.federated", "--help"])
    assert r.returncode == 0
    for flag in ["--data", "--num_agents", "--agent_frac", "--num_corrupt", "--rounds", "--aggr", "--local_ep", "--bs", "--client_lr",
                 "--client_moment", "--server_lr", "--base_class", "--target_class", "--poison_frac", "--pattern_type",
                 "--robustLR_threshold", "--clip", "--noise", "--top_frac", "--snap", "--device", "--num_workers"]:
        assert flag in r.stdout, flag


def test_runner_script_is_valid_bash_and_lists_the_reference_experiments():
    path = os.path.join(ROOT, "scripts", "runner.sh")
    assert subprocess.run(["bash", "-n", path]).returncode == 0
    txt = open(path).read()
    for frag in ("--data=fmnist --local_ep=2 --bs=256 --num_agents=10 --rounds=200", "--num_agents=40", "--num_agents=3383 --agent_frac=0.01",
                 "[fmnist]=4 [cifar10]=8 [fedemnist]=8", "--model=resnet18", "--model=vgg11 --aggr=comed"):
        assert frag in txt, frag
    # dry run: replace the launcher by `echo` and check the 3 x 3 + 2 command lines it generates
    dry = subprocess.run(["bash", "-c", f"sed 's/^  *python /  echo python /' {path} | sed 's/^rm -rf logs.*//' | bash -s 1"],
                         capture_output=True, text=True, cwd=ROOT)
    cmds = [l for l in dry.stdout.splitlines() if l.startswith("python federated.py")]
    assert len(cmds) == 11, dry.stdout + dry.stderr
    assert sum("--robustLR_threshold" in c for c in cmds) == 4 and sum("--data=fedemnist" in c for c in cmds) == 3


def test_bench_reference_arm_reports_unavailable_or_runs_without_gpu():
    r = _run(["bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-1000:]
    assert '"impl": "reference"' in r.stdout


def test_bench_prints_one_json_line_with_the_contract_fields():
    """bench.py's output contract (one JSON line on stdout with the driver's keys), exercised on CPU with a tiny config."""
    import json
    r = _run(["bench.py", "--model", "cnn_mnist", "--data", "fmnist", "--train_size", "256", "--bs", "64", "--steps", "1", "--warmup", "3"],
             timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "clocks", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["metric"] == "fl_rounds_per_sec" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 3 and d["higher_is_better"] is True
    assert d["value"] > 0 and abs(d["value"] - 1000.0 / d["ms_per_step"]) < 1e-6 * d["value"] + 1e-9
    assert d["config"]["model"] == "cnn_mnist" and d["data"] == "synthetic"
