"""The TensorFlow custom-op binding (shim/tf_custom_op.cc + shim/bsl_tf_ops.py) cannot be built against TensorFlow in
this image, so it is held to what CAN be checked: the C++ compiles against a mock of the TF op API with the REAL
include/bsl_b200.h (every C-ABI call type-checks), every registered op has a GPU kernel, every entry point it calls is
declared in the header, the Python half parses, only uses registered ops, and gives every differentiable op a gradient."""
import ast
import re
import shutil
import subprocess
from pathlib import Path

import pytest

from boxsegliver_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent
CC = ROOT / "shim" / "tf_custom_op.cc"
PY = ROOT / "shim" / "bsl_tf_ops.py"


def _snake(name):
    return re.sub(r"(?<=[a-z0-9])(?=[A-Z])|(?<=[A-Z])(?=[A-Z][a-z])", "_", name).lower().replace("2_d", "2d").replace("3_d", "3d").replace("2x2", "2x2")


def test_shim_compiles_against_the_api_mock():
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    r = subprocess.run([gxx, "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", f"-I{ROOT / 'shim' / 'tf_stub'}",
                        f"-I{ROOT / 'include'}", "-DBSL_TF_STUB", str(CC)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_every_op_has_a_kernel_and_only_declared_entry_points_are_called():
    src = CC.read_text()
    ops = re.findall(r'REGISTER_OP\("(\w+)"\)', src)
    kernels = re.findall(r'REGISTER_KERNEL_BUILDER\(Name\("(\w+)"\)', src)
    assert len(ops) >= 20 and sorted(ops) == sorted(kernels)
    called = set(re.findall(r"\b(bsl_[A-Za-z0-9_]+)\s*\(", src)) - {"bsl_tf"}
    declared = set(_lib.exported_symbols_in_header())
    assert called <= declared, called - declared
    # the op families of INTEGRATION.md's table
    for fam in ("bsl_conv2d_fprop", "bsl_conv2d_dgrad", "bsl_conv2d_wgrad", "bsl_conv2d_fprop_stats", "bsl_convT2d_fwd",
                "bsl_convT2d_bwd_data", "bsl_convT2d_bwd_filter", "bsl_conv3d_fprop", "bsl_conv3d_dgrad", "bsl_conv3d_wgrad",
                "bsl_convT3d_fwd", "bsl_norm_finalize", "bsl_norm_apply_pool", "bsl_norm_bwd_reduce", "bsl_norm_bwd_apply",
                "bsl_maxpool2x2_bwd_add", "bsl_wxent_fwd_bwd", "bsl_dice_fwd_bwd", "bsl_softmax_threshold", "bsl_adam_step",
                "bsl_momentum_step", "bsl_allreduce_sum_f32", "bsl_fc_fwd", "bsl_norm_modulate", "bsl_conv2d_head_fprop"):
        assert fam in called, fam
    assert "is_training: bool" in src and 'HostMemory("is_training")' in src     # runtime input, not an attribute


def test_python_half_parses_and_matches_the_registered_ops():
    tree = ast.parse(PY.read_text())
    src = CC.read_text()
    ops = set(re.findall(r'REGISTER_OP\("(\w+)"\)', src))
    snake = {_snake(o): o for o in ops}
    used = {n.attr for n in ast.walk(tree) if isinstance(n, ast.Attribute) and isinstance(n.value, ast.Name)
            and n.value.id == "_lib" and n.attr.startswith("bsl_")}
    assert used <= set(snake), used - set(snake)
    grads = set()
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and \
                node.func.attr in ("RegisterGradient", "NotDifferentiable") and node.args:
            grads.add(node.args[0].value)
    assert grads <= ops, grads - ops
    differentiable = {"BslConv2D", "BslConv2DStats", "BslHeadConv", "BslConv2DTranspose", "BslConv3D", "BslConv3DTranspose",
                      "BslNormRelu", "BslWeightedXent", "BslDiceLoss"}
    assert differentiable <= grads
