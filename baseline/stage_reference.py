"""Stage the UNMODIFIED reference files of the hot path under baseline/_ref/ (git-ignored, but NOT
gpurun-ignored, so the staged copy travels to the GPU box with the snapshot — /root/reference does not).

    python baseline/stage_reference.py            # copies from /root/reference (STREAMFORMER_REF overrides)

The reference (Go2Heart/StreamFormer) is plain Python/PyTorch and not pip-installable (no setup.py; its
pyproject.toml only configures black/isort), so "installing" it means placing the four files the path
lives in where `bench.py --impl reference` can import them:

    models/__init__.py, models/configuration_streamformer.py, models/modeling_timesformer_siglip.py
        -> baseline/_ref/models/                        (TimesformerMultiTaskingModelSigLIP, root copy)
    downstream/VideoQA/llava/model/multimodal_encoder/timesformer_encoder.py
        -> baseline/_ref/kv_twin/timesformer_encoder.py (the KV-cache twin)

Nothing is edited; baseline/_ref/STAGED.json records the sha256 of every source so the staged copy
can be checked against the mount.  The reference sources never enter the git history.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = [
    ("models/__init__.py", "models/__init__.py"),
    ("models/configuration_streamformer.py", "models/configuration_streamformer.py"),
    ("models/modeling_timesformer_siglip.py", "models/modeling_timesformer_siglip.py"),
    ("downstream/VideoQA/llava/model/multimodal_encoder/timesformer_encoder.py", "kv_twin/timesformer_encoder.py"),
]


def stage(ref_root: str | None = None, quiet: bool = False) -> bool:
    """Returns True when baseline/_ref holds the reference afterwards (freshly staged or already there)."""
    ref_root = ref_root or os.environ.get("STREAMFORMER_REF", "/root/reference")
    if not os.path.isdir(ref_root):
        ok = staged()
        if not quiet:
            print(f"{ref_root} not mounted; baseline/_ref {'already staged' if ok else 'ABSENT'}")
        return ok
    manifest = {}
    for src, dst in FILES:
        s, d = os.path.join(ref_root, src), os.path.join(DEST, dst)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest[dst] = {"source": src, "sha256": hashlib.sha256(open(s, "rb").read()).hexdigest()}
    with open(os.path.join(DEST, "STAGED.json"), "w") as f:
        json.dump({"reference": "Go2Heart/StreamFormer", "files": manifest}, f, indent=1)
    if not quiet:
        print(f"staged {len(FILES)} reference files under {DEST}")
    return True


def staged() -> bool:
    return all(os.path.exists(os.path.join(DEST, d)) for _, d in FILES)


def import_reference():
    """(StreamformerConfig, TimesformerMultiTaskingModelSigLIP) of the staged, unmodified reference."""
    if not staged():
        raise ImportError("baseline/_ref is not staged: run python baseline/stage_reference.py where /root/reference is mounted")
    if DEST not in sys.path:
        sys.path.insert(0, DEST)
    import importlib
    for name in [m for m in sys.modules if m == "models" or m.startswith("models.")]:
        mod = sys.modules[name]
        if not getattr(mod, "__file__", "").startswith(DEST):
            del sys.modules[name]
    models = importlib.import_module("models")
    return models.StreamformerConfig, models.TimesformerMultiTaskingModelSigLIP


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
