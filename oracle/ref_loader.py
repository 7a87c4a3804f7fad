"""TEST INFRASTRUCTURE ONLY -- loader that EXECUTES the reference's own code objects.

Used only in the build container (where /root/reference exists) by
oracle/make_golden.py and by the `not gpu` tests that pin oracle/idff_oracle.py to
the reference.  Nothing here travels into the product path, and nothing here is
needed on the GPU box (the golden vectors under tests/golden/ are).

The reference package cannot be imported as a package in this image
(`torchcfm/__init__.py:1` -> `optimal_transport.py:7` needs POT; the example
scripts run training/plots at import and need torchdiffeq/torchsde/matplotlib), so:
  * `torchcfm/models/models.py` and `torchcfm/conditional_flow_matching[_dynamic].py`
    are loaded file-by-file under a package shim with a stub `ot` module;
  * classes/functions defined inside scripts are pulled out by name with `ast`,
    compiled and executed unmodified in a namespace that holds torch/numpy/math.
No reference source is copied: the code objects are built from the files where they lie.
"""
from __future__ import annotations

import ast
import importlib.util
import math
import os
import sys
import types

import numpy as np
import torch

REF_ROOT = os.environ.get("IDFF_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "torchcfm"))


def _ensure_pkg_shim():
    if "ot" not in sys.modules:
        ot = types.ModuleType("ot")
        ot.emd = None
        ot.sinkhorn = None
        ot.unbalanced = types.SimpleNamespace(sinkhorn_knopp_unbalanced=None)
        ot.partial = types.SimpleNamespace(entropic_partial_wasserstein=None)
        sys.modules["ot"] = ot
    if "torchcfm" not in sys.modules or not getattr(sys.modules["torchcfm"], "_idff_shim", False):
        pkg = types.ModuleType("torchcfm")
        pkg.__path__ = [os.path.join(REF_ROOT, "torchcfm")]
        pkg._idff_shim = True
        sys.modules["torchcfm"] = pkg
        sub = types.ModuleType("torchcfm.models")
        sub.__path__ = [os.path.join(REF_ROOT, "torchcfm", "models")]
        sys.modules["torchcfm.models"] = sub


def _load_file(dotted: str, relpath: str):
    _ensure_pkg_shim()
    if dotted in sys.modules:
        return sys.modules[dotted]
    spec = importlib.util.spec_from_file_location(dotted, os.path.join(REF_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[dotted] = mod
    spec.loader.exec_module(mod)
    return mod


def models():
    """reference torchcfm/models/models.py (MLP :4-21, MLP_Embedding :23-42)."""
    return _load_file("torchcfm.models.models", "torchcfm/models/models.py")


def cfm(dynamic: bool = False):
    """reference torchcfm/conditional_flow_matching[_dynamic].py."""
    _load_file("torchcfm.optimal_transport", "torchcfm/optimal_transport.py")
    if dynamic:
        return _load_file("torchcfm.conditional_flow_matching_dynamic",
                          "torchcfm/conditional_flow_matching_dynamic.py")
    return _load_file("torchcfm.conditional_flow_matching", "torchcfm/conditional_flow_matching.py")


def script_symbols(relpath: str, names, extra_ns=None):
    """Compile + exec the named top-level ClassDef/FunctionDef nodes of a reference script."""
    path = os.path.join(REF_ROOT, relpath)
    with open(path, "r") as fh:
        tree = ast.parse(fh.read(), filename=path)
    ns = {"torch": torch, "np": np, "math": math, "__name__": "idff_ref_slice"}
    if extra_ns:
        ns.update(extra_ns)
    want = set(names)
    for node in tree.body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and node.name in want:
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, path, "exec"), ns)
            want.discard(node.name)
    if want:
        raise KeyError(f"{relpath}: symbols not found: {sorted(want)}")
    return {n: ns[n] for n in names}


def script_method(relpath: str, class_name: str, method_name: str, extra_ns=None):
    """Compile + exec ONE method of a class defined in a reference script as a plain function (first argument `self`):
    for classes whose constructor cannot run here (TrajectoryGenerator loads an unshipped dataset)."""
    path = os.path.join(REF_ROOT, relpath)
    with open(path, "r") as fh:
        tree = ast.parse(fh.read(), filename=path)
    ns = {"torch": torch, "np": np, "math": math, "__name__": "idff_ref_slice"}
    if extra_ns:
        ns.update(extra_ns)
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == class_name:
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name == method_name:
                    exec(compile(ast.Module(body=[sub], type_ignores=[]), path, "exec"), ns)
                    return ns[method_name]
    raise KeyError(f"{relpath}: {class_name}.{method_name} not found")


ORDER2 = "2D_examples/Autograd_example_order2.py"
ORDER3 = "2D_examples/Autograd_example_order3.py"
ORDER1 = "2D_examples/example_order1_with_prediction_head.py"
MD = "timeseris_example/MD_simulation.py"


def load_2d_checkpoint(which: str):
    """state_dict of 2D_examples/sample_{which}/best_IDFF_model.pt into the reference MLP."""
    MLP = models().MLP
    net = MLP(dim=2, w=256, time_varying=True)
    sd = torch.load(os.path.join(REF_ROOT, "2D_examples", f"sample_{which}", "best_IDFF_model.pt"),
                    map_location="cpu", weights_only=True)
    net.load_state_dict(sd)
    return net.eval()


def load_md_checkpoint():
    """whole-module pickle timeseris_example/MD/polyALA/MD_trained_polyALA.pt (MD_simulation.py:148,556)."""
    import collections
    m = models()
    torch.serialization.add_safe_globals([
        m.MLP_Embedding, torch.nn.Embedding, torch.nn.Sequential, torch.nn.Linear, torch.nn.SELU,
        set, collections.OrderedDict,
    ])
    net = torch.load(os.path.join(REF_ROOT, "timeseris_example/MD/polyALA/MD_trained_polyALA.pt"),
                     map_location="cpu", weights_only=True)
    return net.eval()
