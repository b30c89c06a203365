"""Recipe for ``oracle/_ref/`` (TEST / BASELINE INFRASTRUCTURE, never imported by the product package).

The reference is Python: "building" it means staging the few source files of the hot path that import cleanly on
their own, UNMODIFIED, from where they lie under ``/root/reference`` into ``oracle/_ref/`` (git-ignored, so no
reference source enters the history; not gpurun-ignored, so the files travel to the GPU box where ``/root/reference``
does not exist).  They serve two purposes only:

* ``bench.py --impl reference`` / ``cpu_baseline``: the CPU arm times the reference's OWN ``SwinTransformerV2`` module
  (mvuld/models/swin_transformer_v2.py:503-652) and ``Rs_GCN`` module (mvuld/models/Rs_GCN.py:7-73) instead of the
  oracle's restatement of them;
* ``tools/make_golden.py``: the golden vectors under ``tests/golden`` are produced by the same modules.

What cannot be staged: ``GraphModel.py`` (imports dgl), ``unixcoder.py`` (imports the pinned transformers 4.18 API and
downloads weights), ``main_bigvul.py`` (dgl, timm, yacs, torchmetrics, a dataset) -- SURVEY.md section 8c.  The only
third-party symbols the two staged files need are three names of ``timm.models.layers``; ``shim_timm`` provides them
(DropPath is the identity in eval mode, which is the mode every parity and throughput run uses).
"""
from __future__ import annotations

import importlib.util
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"
OUT = os.path.join(HERE, "_ref")
FILES = {"mvuld/models/swin_transformer_v2.py": "swin_transformer_v2.py", "mvuld/models/Rs_GCN.py": "Rs_GCN.py"}


def build_ref(verbose: bool = True) -> bool:
    """Stage the reference files (only where /root/reference exists, i.e. in the build container)."""
    if not os.path.isdir(REF_ROOT):
        return available()
    os.makedirs(OUT, exist_ok=True)
    for rel, name in FILES.items():
        shutil.copyfile(os.path.join(REF_ROOT, rel), os.path.join(OUT, name))
    with open(os.path.join(OUT, "README"), "w") as fh:
        fh.write("Unmodified copies of reference files staged by oracle/ref_build.py (git-ignored; CPU baseline only).\n")
    if verbose:
        print(f"[oracle/_ref] staged {sorted(FILES.values())}")
    return True


def available() -> bool:
    return all(os.path.exists(os.path.join(OUT, n)) for n in FILES.values())


def shim_timm():
    """The three ``timm.models.layers`` symbols swin_transformer_v2.py:11 imports (timm is not installed here)."""
    if "timm.models.layers" in sys.modules:
        return
    import torch.nn as nn

    class DropPath(nn.Module):
        def __init__(self, p=0.0):
            super().__init__()
            self.p = p

        def forward(self, x):
            assert not self.training, "shim DropPath is eval-only"
            return x

    layers = types.ModuleType("timm.models.layers")
    layers.DropPath = DropPath
    layers.to_2tuple = lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x)
    layers.trunc_normal_ = nn.init.trunc_normal_
    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    timm.models, models.layers = models, layers
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})


def load(name: str):
    """Import a staged reference file as a module (``swin_transformer_v2`` / ``Rs_GCN``)."""
    path = os.path.join(OUT, name + ".py")
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path}: run oracle/ref_build.py where /root/reference exists")
    shim_timm()
    spec = importlib.util.spec_from_file_location("mvuld_ref_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print("available" if build_ref() else "reference not present and nothing staged")
