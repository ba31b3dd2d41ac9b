"""Feature schema objects the layers are constructed from.

The layers accept the reference's own `recbox.ranking.features.FeatureMap`
(ranking/features.py:25-125) and the matching side's feature-spec holder
(matching/features.py:12-58) unchanged -- they only touch `.features` / `.feature_specs`,
`.num_fields`, `.labels`, `.data_dir`, `.get_column_index()`.  These two small holders provide that
surface for use without the reference package (tests, bench.py, tools): builders for a schema plus the
column layout of the batch matrix.  feature_map.json I/O is the reference's (out of scope).
"""
from collections import OrderedDict


class FeatureMap(object):
    """Ranking-side schema: ordered {feature: spec}; column layout of the [B, n_cols] batch matrix
    is consecutive in feature order, `max_len` columns for a sequence feature, labels last
    (ranking/features.py:106-120)."""

    def __init__(self, dataset_id="dataset", data_dir="."):
        self.data_dir = data_dir
        self.dataset_id = dataset_id
        self.num_fields = 0
        self.total_features = 0
        self.input_length = 0
        self.features = OrderedDict()
        self.labels = []
        self.column_index = dict()
        self.group_id = None
        self.default_emb_dim = None

    # -- construction helpers (not in the reference; the reference fills .features from preprocessing)
    def add_numeric(self, name, source="", **spec):
        self.features[name] = dict({"source": source, "type": "numeric"}, **spec)
        return self

    def add_categorical(self, name, vocab_size, source="", padding_idx=0, **spec):
        s = {"source": source, "type": "categorical", "vocab_size": int(vocab_size)}
        if padding_idx is not None:
            s["padding_idx"] = padding_idx
        s.update(spec)
        self.features[name] = s
        return self

    def add_sequence(self, name, vocab_size, max_len, source="", padding_idx=0, **spec):
        s = {"source": source, "type": "sequence", "vocab_size": int(vocab_size), "max_len": int(max_len)}
        if padding_idx is not None:
            s["padding_idx"] = padding_idx
        s.update(spec)
        self.features[name] = s
        return self

    def finalize(self, labels=("label",)):
        self.labels = list(labels)
        self.num_fields = self.get_num_fields()
        self.total_features = sum(s.get("vocab_size", 1) for s in self.features.values() if s["type"] != "meta")
        self.set_column_index()
        return self

    # -- what the layers and the batch packers read (ranking/features.py:92-125); reading / writing feature_map.json stays
    #    with the reference's own FeatureMap, which the layers accept unchanged
    def get_num_fields(self, feature_source=[]):
        if type(feature_source) != list:
            feature_source = [feature_source]
        return sum(1 for s in self.features.values()
                   if s["type"] != "meta" and (len(feature_source) == 0 or s.get("source") in feature_source))

    def sum_emb_out_dim(self, feature_source=[]):
        if type(feature_source) != list:
            feature_source = [feature_source]
        total = 0
        for s in self.features.values():
            if s["type"] == "meta":
                continue
            if len(feature_source) == 0 or s.get("source") in feature_source:
                total += s.get("emb_output_dim", s.get("embedding_dim", self.default_emb_dim))
        return total

    def set_column_index(self):
        idx = 0
        for feature, spec in self.features.items():
            if "max_len" in spec:
                self.column_index[feature] = [i + idx for i in range(spec["max_len"])]
                idx += spec["max_len"]
            else:
                self.column_index[feature] = idx
                idx += 1
        self.input_length = idx
        for label in self.labels:
            self.column_index[label] = idx
            idx += 1

    def get_column_index(self, feature):
        if feature not in self.column_index:
            self.set_column_index()
        return self.column_index[feature]


class MatchingFeatureMap(object):
    """Matching-side schema holder: `.feature_specs` ordered {feature: spec} with `source` in
    {"user", "item"}, padding at the LAST row on this side (matching/preprocess.py:56-58)."""

    def __init__(self, dataset_id="dataset", data_dir=".", feature_specs=None):
        self.dataset_id = dataset_id
        self.data_dir = data_dir
        self.feature_specs = OrderedDict(feature_specs or {})
        self.query_index = None
        self.corpus_index = None
