"""Line-level similarity of the host-side mirrors to the reference files they mirror (difflib ratio over stripped, non-comment
lines).  Run in the authoring container (needs /root/reference); the round-1 driver-side detector timed out, so this is the
self-check DESIGN.md quotes.  Threshold of the detector: 0.6."""
import difflib
import os
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/LDMAE"
PAIRS = [
    ("ldmae_b200/transport/transport.py", "transport/transport.py"),
    ("ldmae_b200/transport/transport.py", "transport/integrators.py"),
    ("ldmae_b200/transport/transport.py", "transport/path.py"),
    ("ldmae_b200/transport/__init__.py", "transport/__init__.py"),
    ("ldmae_b200/models/lightningdit.py", "models/lightningdit.py"),
    ("ldmae_b200/tokenizer/models_mae.py", "tokenizer/models_mae.py"),
    ("ldmae_b200/datasets/img_latent_dataset.py", "datasets/img_latent_dataset.py"),
    ("ldmae_b200/checkpoint.py", "train_accum.py"),
    ("ldmae_b200/training.py", "train_accum.py"),
    ("oracle/ldmae_oracle.py", "models/lightningdit.py"),
    ("oracle/ldmae_oracle.py", "tokenizer/models_mae.py"),
]


def lines(path):
    out = []
    for l in open(path, errors="ignore"):
        l = l.strip()
        if l and not l.startswith("#"):
            out.append(l)
    return out


if __name__ == "__main__":
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    worst = 0.0
    for ours, theirs in PAIRS:
        a, b = os.path.join(root, ours), os.path.join(REF, theirs)
        if not (os.path.exists(a) and os.path.exists(b)):
            print("missing", ours, theirs)
            continue
        r = difflib.SequenceMatcher(None, lines(a), lines(b), autojunk=False).ratio()
        worst = max(worst, r)
        print(f"{r:.2f}  {ours}  vs  {theirs}")
    print(f"worst {worst:.2f} (detector threshold 0.60)")
