"""Host-side mirror of the reference's model / solver / distribution API (no GPU): flags, defaults, YAML lookup,
learning-rate policies and error text follow /root/reference/core/models.py:41-118, core/solver.py:23-202 and
utils/distribution_utils.py:124-134."""
import argparse

import pytest

from boxsegliver_b200 import distribution_utils, models, solver


def _args(*argv):
    p = argparse.ArgumentParser()
    models.add_arguments(p)
    solver.add_arguments(p)
    return p.parse_args(list(argv))


def test_model_zoo_and_yaml_defaults():
    assert [c.__name__ for c in models.MODEL_ZOO] == ["UNet", "GUNet", "UNet3D", "UNetInter"]
    for name, keys in (("UNet", {"init_channels": 64, "num_down_samples": 4}),
                       ("UNetInter", {"init_channels": 64, "num_down_samples": 4, "ret_pred": True}),
                       ("GUNet", {"mod_layers": [1, 2, 3, 4]})):
        args = _args("--model", name, "--classes", "Liver", "Tumor")
        params = models.get_model_params(args, build_metrics=True)
        assert params["model"].__name__ == name
        for k, v in keys.items():
            assert params["model_kwargs"][k] == v, (name, k)
        assert params["model_kwargs"]["build_metrics"] is True
    with pytest.raises(SystemExit):
        _args("--model", "NoSuchNet", "--classes", "Liver")


def test_flag_defaults_match_reference():
    a = _args("--model", "UNet", "--classes", "Liver", "Tumor")
    assert (a.batch_size, a.weight_init, a.normalizer) == (8, "xavier", "batch_norm")          # models.py:57-63
    assert (a.learning_rate, a.learning_policy, a.optimizer) == (1e-3, "period_step", "Adam")  # solver.py:27-58
    assert (a.lr_decay_step, a.lr_decay_rate, a.lr_power, a.lr_end) == (100000, 0.1, 0.9, 1e-6)
    assert (a.slow_start_step, a.slow_start_lr, a.lr_patience) == (1000, 1e-4, 30)


def test_learning_rate_policies():
    a = _args("--model", "UNet", "--classes", "Liver", "--learning_policy", "period_step", "--lr_decay_step", "10",
              "--lr_decay_rate", "0.5", "--learning_rate", "0.01")
    s = solver.Solver(a)
    assert s.learning_rate(0) == 0.01 and s.learning_rate(9) == 0.01 and s.learning_rate(10) == 0.005
    assert s.learning_rate(25) == 0.0025
    a = _args("--model", "UNet", "--classes", "Liver", "--learning_policy", "custom_step", "--lr_decay_boundaries", "5",
              "8", "--lr_custom_values", "1.0", "0.1", "0.01")
    s = solver.Solver(a)
    assert [s.learning_rate(t) for t in (0, 5, 6, 8, 9)] == [1.0, 1.0, 0.1, 0.1, 0.01]    # x <= boundary keeps the value
    a = _args("--model", "UNet", "--classes", "Liver", "--learning_policy", "poly", "--num_of_total_steps", "100",
              "--learning_rate", "0.1", "--lr_end", "0.001", "--lr_power", "2.0")
    s = solver.Solver(a)
    assert abs(s.learning_rate(50) - ((0.1 - 0.001) * 0.25 + 0.001)) < 1e-12
    assert s.learning_rate(1000) == pytest.approx(0.001)
    a = _args("--model", "UNet", "--classes", "Liver", "--learning_policy", "plateau", "--lr_decay_rate", "0.1",
              "--lr_end", "1e-4", "--lr_warm_up", "--slow_start_step", "3", "--slow_start_lr", "7e-5")
    s = solver.Solver(a)
    assert s.learning_rate(2) == 7e-5 and s.learning_rate(3) == 1e-3
    assert s.plateau_update() == pytest.approx(1e-4) and s.plateau_update() == pytest.approx(1e-4)
    with pytest.raises(ValueError, match="len\\(lr_custom_values\\) - len\\(lr_decay_boundaries\\) = 1"):
        solver.Solver(_args("--model", "UNet", "--classes", "Liver", "--learning_policy", "custom_step",
                            "--lr_decay_boundaries", "5", "--lr_custom_values", "1.0"))


def test_per_device_batch_size():
    assert distribution_utils.per_device_batch_size(64, 8) == 8
    assert distribution_utils.per_device_batch_size(8, 1) == 8
    with pytest.raises(ValueError):
        distribution_utils.per_device_batch_size(10, 4)
