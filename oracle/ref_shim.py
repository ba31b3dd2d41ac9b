"""Import shims that make the UNMODIFIED reference (reczoo/RecBox, mounted read-only at
/root/reference) importable on CPU in the dev container.

TEST INFRASTRUCTURE ONLY.  Used by oracle/make_golden.py (to mint tests/golden/*.npz) and by the
container-only cross-check tests.  /root/reference does not exist on the GPU box, so nothing on the
product path, in `-m gpu` tests, smoke() or bench.py imports this module.

Why shims are needed:
  * h5py / faiss are imported at module top by the reference but only used for pretrained-embedding
    I/O and ANN evaluation (recbox/core/pytorch/layers/embedding.py:3,
    recbox/ranking/pytorch/layers/embeddings/feature_embedding.py:20, recbox/utils/ann/faiss.py:1).
  * recbox/ranking is a rename of FuxiCTR v2 and still imports `fuxictr.*`
    (feature_embedding.py:24-25, ranking_model.py:23-25); the modules it wants are its own siblings.
  * rechub's SASRec imports `torch_rechub.*` (third_party/rechub/models/matching/sasrec.py:13-14).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("RECBOX_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "recbox"))


_done = False


def install():
    """Idempotently register the aliases; returns the `recbox.ranking.pytorch.layers` module."""
    global _done
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if not _done:
        for name in ("h5py", "faiss"):
            sys.modules.setdefault(name, types.ModuleType(name))
        # `import recbox` pulls five vendored zoos (and TF); make the package a bare namespace.
        for pkg in ("recbox", "recbox.third_party"):
            if pkg not in sys.modules:
                m = types.ModuleType(pkg)
                m.__path__ = [os.path.join(REFERENCE_ROOT, *pkg.split("."))]
                sys.modules[pkg] = m
        fx = types.ModuleType("fuxictr")
        fx.__path__ = []
        sys.modules["fuxictr"] = fx
        import recbox.ranking.features, recbox.ranking.metrics, recbox.ranking.utils  # noqa
        sys.modules["fuxictr.features"] = sys.modules["recbox.ranking.features"]
        sys.modules["fuxictr.metrics"] = sys.modules["recbox.ranking.metrics"]
        sys.modules["fuxictr.utils"] = sys.modules["recbox.ranking.utils"]
        import recbox.ranking.pytorch as rkp
        sys.modules["fuxictr.pytorch"] = rkp
        import recbox.ranking.pytorch.torch_utils as tu
        sys.modules["fuxictr.pytorch.torch_utils"] = tu
        import recbox.ranking.pytorch.layers as L
        sys.modules["fuxictr.pytorch.layers"] = L
        _done = True
    return sys.modules["recbox.ranking.pytorch.layers"]


def install_rechub():
    install()
    import recbox.third_party.rechub as rh
    import recbox.third_party.rechub.basic as rb
    import recbox.third_party.rechub.basic.features as rf
    import recbox.third_party.rechub.basic.layers as rl
    sys.modules["torch_rechub"] = rh
    sys.modules["torch_rechub.basic"] = rb
    sys.modules["torch_rechub.basic.features"] = rf
    sys.modules["torch_rechub.basic.layers"] = rl
    return rh
