"""Import the UNMODIFIED IDEAL-NeRF reference from /root/reference on a CPU-only host.

TEST INFRASTRUCTURE ONLY.  This file is used by ``tests/golden/make_golden.py`` (run once, in the
build container where /root/reference exists) to produce the committed golden fixtures.  It is
never imported by the product package, by ``-m gpu`` tests, by ``smoke()`` or by ``bench.py``.

The reference parses ``sys.argv`` and touches CUDA at import time
(NeRFs/HeadNeRF/helper.py:141-142,191; NeRFs/HeadNeRF/train/audio_exp_nerf.py:25-36), and imports
four third-party modules that are absent from this image.  The recipe below stubs
those modules, supplies the flags, and turns ``Tensor.cuda`` into the identity so the reference's
own code runs on the CPU in fp32.
"""
import importlib
import os
import sys
import tempfile
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    """$IDEAL_NERF_REFERENCE, else /root/reference (build container), else oracle/_ref (the copy oracle/make_ref.py made; GPU box)."""
    for p in (os.environ.get("IDEAL_NERF_REFERENCE"), "/root/reference", os.path.join(_HERE, "_ref")):
        if p and os.path.isdir(os.path.join(p, "NeRFs")):
            return p
    return "/root/reference"


REF_ROOT = _find_root()


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "NeRFs"))

NEAR = 0.5772005200386048   # NeRFs/HeadNeRF/configs/audio_expr_nerf/may/paper_model/torso_bg.txt:11
FAR = 1.1772005200386046    # ...:12


def _stub_modules():
    for name in ("face_alignment", "imageio"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    if "natsort" not in sys.modules:
        m = types.ModuleType("natsort")
        m.natsorted = sorted
        sys.modules["natsort"] = m
    if "configargparse" not in sys.modules:
        import argparse
        m = types.ModuleType("configargparse")

        class ArgumentParser(argparse.ArgumentParser):
            def add_argument(self, *a, **k):
                k.pop("is_config_file", None)
                return super().add_argument(*a, **k)

        m.ArgumentParser = ArgumentParser
        sys.modules["configargparse"] = m


def import_head(perturb=0.0, n_samples=64, n_importance=128, dim_aud=64, dim_expr=76,
                near=NEAR, far=FAR, force_cpu=False):
    """Return the reference module NeRFs.HeadNeRF.train.audio_exp_nerf (class-form renderer).  force_cpu: keep the reference's
    hard-coded `.cuda()` calls (helper.py:191,197,204,279,282) on the host even when a GPU is visible (CPU baseline on the GPU box);
    without it, on a GPU box, call torch.set_default_tensor_type('torch.cuda.FloatTensor') first like the reference's __main__ (:598)."""
    import torch
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    _stub_modules()
    if force_cpu or not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    vis = tempfile.mkdtemp(prefix="inerf_vis_")
    sys.argv = ["ref", "--N_samples", str(n_samples), "--N_importance", str(n_importance),
                "--dim_aud", str(dim_aud), "--dim_expr", str(dim_expr), "--perturb", str(perturb),
                "--near", repr(near), "--far", repr(far), "--vis_path", vis]
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    mod = importlib.import_module("NeRFs.HeadNeRF.train.audio_exp_nerf")
    torch.autograd.set_detect_anomaly(False)
    return mod


def import_torso_helpers():
    """NeRFs/TorsoNeRF/run_nerf_helpers.py is device-parametrised and imports cleanly."""
    import torch
    _stub_modules()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    mod = importlib.import_module("NeRFs.TorsoNeRF.run_nerf_helpers")
    torch.autograd.set_detect_anomaly(False)   # run_nerf_helpers.py:7 switches it on
    return mod
