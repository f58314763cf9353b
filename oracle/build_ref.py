"""ORACLE tooling — test / baseline infrastructure, not product code.

Recipe that makes the reference's OWN modules for the hot path runnable beside our code on the GPU
box: it copies the few pure-Python files of the path, unmodified, from where they lie under
/root/reference into `oracle/_ref/refsrc/` (git-ignored: reference sources never enter this
repository's history; not gpurun-ignored: the directory travels to the GPU box like a built .so).

    python oracle/build_ref.py        # also run by __graft_entry__.build() when /root/reference exists

Files (all stock-PyTorch Python, no build step):
    src/ncf/models.py               NCF                      (forward / init / load_pretrain_weights)
    src/training/metrics.py         metrics                  (leave-one-out HR / NDCG)
    src/distillation/base.py        BaseDistillation
    src/distillation/response.py    ResponseDistillation, SoftTargetDistillation
    src/distillation/feature.py     FeatureDistillation
    src/distillation/attention.py   AttentionDistillation
The package __init__ files are written empty here (the reference's own `src/data/__init__.py` pulls
matplotlib in, which this image does not have; none of the files above needs it).

Consumers: bench.py `--impl reference` / `cpu_baseline` / `gpu_eager_reference` (kind "reference"
when this directory exists, else the port in oracle/torch_port.py), and tests that pin the oracle.
Nothing under ncf_b200/ or scripts/ may import it.
"""
from __future__ import annotations

import shutil
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference")
OUT = HERE / "_ref" / "refsrc"
FILES = ["src/ncf/models.py", "src/training/metrics.py", "src/distillation/base.py",
         "src/distillation/response.py", "src/distillation/feature.py", "src/distillation/attention.py"]


def build(verbose: bool = True) -> bool:
    if not REF.exists():
        if verbose:
            print(f"{REF} not present: oracle/_ref not rebuilt (a prebuilt copy, if any, is used as is)")
        return (OUT / "src" / "ncf" / "models.py").exists()
    if OUT.exists():
        shutil.rmtree(OUT)
    for rel in FILES:
        dst = OUT / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(REF / rel, dst)
    for d in [OUT / "src", OUT / "src/ncf", OUT / "src/training", OUT / "src/distillation"]:
        (d / "__init__.py").write_text("")
    (OUT.parent / "README").write_text(
        "Unmodified copies of reference files, written by oracle/build_ref.py. Git-ignored. Do not edit.\n")
    if verbose:
        print(f"oracle/_ref: {len(FILES)} reference files under {OUT}")
    return True


def load():
    """Imports the vendored reference modules (namespace `refsrc_pkg`), or returns None when
    oracle/_ref has not been built.  The reference's package is called `src`; it is imported under
    that name from oracle/_ref/refsrc with the path inserted only for the duration of the import."""
    if not (OUT / "src" / "ncf" / "models.py").exists():
        return None
    import importlib
    import types
    saved = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, str(OUT))
    try:
        ns = types.SimpleNamespace()
        ns.models = importlib.import_module("src.ncf.models")
        ns.metrics = importlib.import_module("src.training.metrics")
        ns.base = importlib.import_module("src.distillation.base")
        ns.response = importlib.import_module("src.distillation.response")
        ns.feature = importlib.import_module("src.distillation.feature")
        ns.attention = importlib.import_module("src.distillation.attention")
        ns.NCF = ns.models.NCF
        ns.path = str(OUT)
    finally:
        sys.path.remove(str(OUT))
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    return ns


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
