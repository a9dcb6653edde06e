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


def test_optimizer_flags_follow_get_solver_params():
    """core/solver.py:84-97,204-219: no flag -> the reference's defaults; ANY flag -> the flag dict replaces them (so
    TensorFlow's own defaults fill the rest); keywords the chosen optimizer's constructor lacks are TypeErrors."""
    base = ("--model", "UNet", "--classes", "Liver")
    k = solver.engine_optimizer_kwargs
    assert k(_args(*base)) == dict(optimizer="adam", adam_beta1=0.9, adam_beta2=0.99, adam_eps=1e-8)
    assert k(_args(*base, "--adam_beta1", "0.5")) == dict(optimizer="adam", adam_beta1=0.5, adam_beta2=0.999, adam_eps=1e-8)
    assert k(_args(*base, "--adam_beta2", "0.9", "--adam_eps", "1e-4")) == dict(optimizer="adam", adam_beta1=0.9,
                                                                               adam_beta2=0.9, adam_eps=1e-4)
    assert k(_args(*base, "--optimizer", "Momentum")) == dict(optimizer="momentum", momentum=0.9, use_nesterov=False)
    assert k(_args(*base, "--optimizer", "Momentum", "--mm_mm", "0.8", "--mm_nesterov")) == dict(
        optimizer="momentum", momentum=0.8, use_nesterov=True)
    a = _args(*base, "--optimizer", "AdamW")
    a.weight_decay_rate = 3e-5
    assert k(a) == dict(optimizer="adamw", adam_beta1=0.9, adam_beta2=0.99, adam_eps=1e-8, adamw_weight_decay=3e-5)
    with pytest.raises(TypeError, match="missing 1 required positional argument: 'momentum'"):
        k(_args(*base, "--optimizer", "Momentum", "--mm_nesterov"))       # MomentumOptimizer(lr, use_nesterov=True)
    with pytest.raises(TypeError, match="unexpected keyword argument 'momentum'"):
        k(_args(*base, "--mm_mm", "0.5"))                                   # AdamOptimizer(lr, momentum=0.5)
    with pytest.raises(TypeError, match="unexpected keyword argument 'beta1'"):
        k(_args(*base, "--optimizer", "Momentum", "--adam_beta1", "0.5", "--mm_mm", "0.9"))
    with pytest.raises(TypeError, match="weight_decay"):
        k(_args(*base, "--optimizer", "AdamW", "--adam_beta1", "0.5"))
    assert solver.Solver(_args(*base, "--adam_beta1", "0.5")).optimizer_params == {"beta1": 0.5}
    assert solver.Solver(_args(*base)).optimizer_params is None


def test_oracle_variable_names_follow_slim_scoping():
    """The oracle's (and, in tests/test_gpu_host_api.py, the engine's) variable list equals the list derived from
    slim's scoping rules alone -- a reference model_dir restores only if every name matches."""
    from oracle import unet_ref as R
    from tests.slim_names import unet_variable_names
    for norm in ("batch_norm", "instance_norm"):
        cfg = R.UNetCfg(height=32, width=32, normalizer=norm)
        got = set(R.init_params(cfg, seed=0))
        assert got == set(unet_variable_names(4, norm)), sorted(got ^ set(unet_variable_names(4, norm)))
    assert "UNet/ED-Bridge/ED-Bridge_2/BatchNorm/moving_variance" in unet_variable_names()
    assert len(unet_variable_names(with_moving=False)) == 64      # 64 trainable variables
