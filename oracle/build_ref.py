"""TEST / BASELINE INFRASTRUCTURE ONLY -- stages the reference's own hot-path files under oracle/_ref/.

    python oracle/build_ref.py            # here (authoring container), where /root/reference exists

The reference (isno0907/ldmae) is pure Python with no setup.py / pyproject.toml, so there is nothing to compile or
pip-install; the files of the path import verbatim through oracle/refshim.py (which restates the five absent
third-party modules).  /root/reference does not exist on the GPU box, so this recipe copies exactly the files the
shimmed import touches into oracle/_ref/LDMAE/ -- git-ignored (reference sources never enter the history), NOT
gpurun-ignored (the staged copy travels to the box like a built .so).  `bench.py --impl reference` and the
`cpu_baseline` leg then time the reference's OWN modules on the box's host cores (`kind: "reference"`); without the
staged copy they fall back to the oracle port (`kind: "port"`).  The product (ldmae_b200/) never imports any of this.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("LDMAE_REFERENCE_SRC", "/root/reference/LDMAE")
DST = os.path.join(HERE, "_ref", "LDMAE")

# the files the shimmed import of models.lightningdit / transport / tokenizer.models_mae loads
FILES = [
    "models/__init__.py", "models/lightningdit.py", "models/rmsnorm.py", "models/swiglu_ffn.py", "models/pos_embed.py",
    "transport/__init__.py", "transport/transport.py", "transport/path.py", "transport/integrators.py", "transport/utils.py",
    "tokenizer/__init__.py", "tokenizer/models_mae.py", "tokenizer/util/pos_embed.py", "tokenizer/util/misc.py",
]


def build(verbose: bool = False) -> bool:
    """Copies FILES from the reference tree; returns False (and leaves any existing staged copy alone) when the
    reference tree is not present (e.g. on the GPU box)."""
    if not os.path.isdir(os.path.join(SRC, "models")):
        return os.path.isdir(os.path.join(DST, "models"))
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.exists(s):
            raise FileNotFoundError(f"reference file missing: {s}")
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not os.path.exists(d) or os.path.getmtime(d) < os.path.getmtime(s) or os.path.getsize(d) != os.path.getsize(s):
            shutil.copyfile(s, d)
            if verbose:
                print("staged", rel)
    return True


def available() -> bool:
    return all(os.path.exists(os.path.join(DST, rel)) for rel in FILES)


if __name__ == "__main__":
    ok = build(verbose=True)
    print("oracle/_ref", "ready" if ok and available() else "NOT available (no reference tree here)")
    sys.exit(0 if ok else 1)
