"""Command surface mirrors the reference's main.py:241-284: group options -a/--algorithm, -m/--model; sub-commands
train | test | record, each with -e/--environment.  No GPU needed: only the argument handling and the loud failure
without CUDA are exercised here."""
import pytest

click = pytest.importorskip("click")
from click.testing import CliRunner  # noqa: E402

from so100_mujoco_rl_b200.cli import cli  # noqa: E402


def test_group_and_subcommands_match_the_reference_surface():
    r = CliRunner().invoke(cli, ["--help"])
    assert r.exit_code == 0
    for token in ("-a, --algorithm", "-m, --model", "train", "test", "record"):
        assert token in r.output
    for sub in ("train", "test", "record"):
        h = CliRunner().invoke(cli, [sub, "--help"])
        assert h.exit_code == 0 and "-e, --environment" in h.output
    t = CliRunner().invoke(cli, ["train", "--help"]).output
    for token in ("--num-envs", "--learner", "--eval-freq", "--save-freq", "--tensorboard-log", "--trainer"):
        assert token in t


def test_environment_is_required_and_unknown_algorithms_are_refused():
    assert CliRunner().invoke(cli, ["train"]).exit_code != 0
    r = CliRunner().invoke(cli, ["-a", "DDPG", "train", "-e", "Env01"], obj={})
    assert r.exit_code != 0 and "native trainer implements PPO" in r.output


def test_train_fails_loudly_without_cuda(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present: the loud-failure path is not reachable")
    r = CliRunner().invoke(cli, ["-a", "PPO", "train", "-e", "Env01", "--num-envs", "8", "--out", str(tmp_path)], obj={})
    assert r.exit_code != 0
    assert isinstance(r.exception, (RuntimeError, ValueError)) and "CUDA" in str(r.exception)   # no CPU fallback


def test_evaluation_needs_a_model():
    r = CliRunner().invoke(cli, ["test", "-e", "Env01"], obj={})
    assert r.exit_code != 0 and "--model is required" in r.output
