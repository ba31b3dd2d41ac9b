"""recbox_b200 -- B200 (sm_100a) implementation of RecBox's embedding + feature-interaction hot path.

Layout:
  csrc/        hand-written CUDA kernels + the C ABI of include/recbox_b200.h (-> librecbox_b200.so)
  _lib.py      in-tree build + ctypes loader (no fallback: missing library = exception)
  ops.py       one-call-per-op tensor wrappers over the C ABI
  layers.py    the reference's nn.Module operator API on top (same names / ctor args / state_dict keys)
  features.py  small FeatureMap schema holders for tests / bench (the layers accept the reference's own FeatureMap unchanged)
  sharded.py   row-sharded table across the GPUs of one box (NCCL all-to-all or NVLink peer access)
  optim.py     exact-dense clip + Adam on the fused tables, and the touched-rows variant
  loader.py    packed input pipeline (ids / dense converted once to pinned int32 / fp32 blocks), epoch negative sampler
  retrieval.py FaissIndex / evaluate_metrics drop-ins on the fused top-k and ranking-metric kernels

`install()` rebinds the reference's own symbols (recbox.ranking.pytorch.layers.*, recbox.core.pytorch.layers.*,
recbox.matching.pytorch.layers.*, and the fuxictr.* aliases RecBox's ranking package still imports) to the
classes of layers.py, so unmodified model code constructs the fused modules.
"""
import sys

__version__ = "0.1.0"

from ._lib import RbxError, build, load  # noqa: F401

# reference module prefix -> {attribute: name in recbox_b200.layers}
_RANKING = {"FeatureEmbedding": "FeatureEmbedding", "FeatureEmbeddingDict": "FeatureEmbeddingDict",
            "InnerProductInteraction": "InnerProductInteraction", "LogisticRegression": "LogisticRegression",
            "FactorizationMachine": "FactorizationMachine", "MaskedAveragePooling": "MaskedAveragePooling",
            "MaskedSumPooling": "MaskedSumPooling", "InteractionMachine": "InteractionMachine", "MLP_Block": "MLP_Block",
            "CrossInteraction": "CrossInteraction", "CrossNet": "CrossNet", "CrossNetV2": "CrossNetV2",
            "CompressedInteractionNet": "CompressedInteractionNet", "ScaledDotProductAttention": "ScaledDotProductAttention",
            "DIN_Attention": "DIN_Attention", "MultiHeadTargetAttention": "MultiHeadTargetAttention"}
_CORE = {"EmbeddingLayer": "EmbeddingLayer", "EmbeddingDictLayer": "EmbeddingDictLayer",
         "MaskedAveragePooling": "CoreMaskedAveragePooling", "MaskedSumPooling": "CoreMaskedSumPooling", "MLP_Layer": "MLP_Layer"}
_TARGETS = (("recbox.ranking.pytorch.layers", _RANKING), ("fuxictr.pytorch.layers", _RANKING),
            ("recbox.core.pytorch.layers", _CORE), ("recbox.matching.pytorch.layers", _CORE))
_saved = []


def install(import_reference=True):
    """Rebind the reference's layer classes to the fused B200 ones in every already-imported (and,
    with import_reference, importable) reference module.  Returns the list of (module, attribute)
    pairs that were rebound.  Idempotent; `uninstall()` restores the originals."""
    import importlib
    from . import blocks, layers as ours
    for name in blocks.__all__:                       # the GEMM-shaped consumers live in blocks.py
        if not hasattr(ours, name):
            setattr(ours, name, getattr(blocks, name))
    done = []
    for prefix, table in _TARGETS:
        if import_reference and prefix not in sys.modules:
            try:
                importlib.import_module(prefix)
            except Exception:
                continue
        for modname, mod in list(sys.modules.items()):
            if mod is None or not (modname == prefix or modname.startswith(prefix + ".")):
                continue
            for attr, mine in table.items():
                cur = mod.__dict__.get(attr)
                new = getattr(ours, mine)
                if cur is None or cur is new or not isinstance(cur, type):
                    continue
                _saved.append((mod, attr, cur))
                setattr(mod, attr, new)
                done.append((modname, attr))
    return done


def uninstall():
    while _saved:
        mod, attr, cur = _saved.pop()
        setattr(mod, attr, cur)
