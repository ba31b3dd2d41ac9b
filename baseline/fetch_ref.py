"""Copies the reference's OWN hot-path modules into the git-ignored baseline/_ref/ so that the GPU box (where
/root/reference does not exist) can time the UNMODIFIED reference implementation of the path on its host cores:

    python baseline/fetch_ref.py        (run in the dev container; __graft_entry__.build() calls it when the reference is there)

What travels: recbox/ranking/{features,metrics,utils}.py, recbox/ranking/pytorch/** (the FuxiCTR-style layer set:
FeatureEmbedding, FactorizationMachine, LogisticRegression, InnerProductInteraction, MLP_Block, RankingModel ...) and
recbox/core/pytorch/** (the matching-side layers and losses) -- pure Python, ~3 k lines.  Nothing is edited; the import
aliases the reference needs (fuxictr.* -> recbox.ranking.*, stub h5py / faiss) are applied at import time by
oracle/ref_shim.py with RECBOX_REFERENCE pointing here.  baseline/_ref/ is in .gitignore (never committed) and NOT in
.gpurunignore (it ships with the snapshot)."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("RECBOX_REFERENCE_SRC", "/root/reference")
TREES = ["recbox/ranking", "recbox/core/pytorch"]
SKIP_DIRS = {"tensorflow", "__pycache__", "preprocess"}


def fetch(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "recbox")):
        if verbose:
            print("fetch_ref: no reference at %s (keeping whatever baseline/_ref already holds)" % SRC)
        return False
    n = 0
    for tree in TREES:
        for root, dirs, files in os.walk(os.path.join(SRC, tree)):
            dirs[:] = [d for d in dirs if d not in SKIP_DIRS]
            rel = os.path.relpath(root, SRC)
            os.makedirs(os.path.join(DEST, rel), exist_ok=True)
            for f in files:
                if f.endswith(".py"):
                    shutil.copyfile(os.path.join(root, f), os.path.join(DEST, rel, f))
                    n += 1
    # parents of the copied trees must be importable packages for `import recbox.core.pytorch.layers`
    for pkg in ("recbox/core",):
        init = os.path.join(SRC, pkg, "__init__.py")
        if os.path.exists(init) and not os.path.exists(os.path.join(DEST, pkg, "__init__.py")):
            open(os.path.join(DEST, pkg, "__init__.py"), "w").close()     # bare package: the reference's own pulls TF twins
    with open(os.path.join(DEST, "SOURCE.txt"), "w") as f:
        f.write("copied unmodified from %s by baseline/fetch_ref.py (%d files)\n" % (SRC, n))
    if verbose:
        print("fetch_ref: %d reference files -> %s" % (n, DEST))
    return True


if __name__ == "__main__":
    sys.exit(0 if fetch() else 1)
