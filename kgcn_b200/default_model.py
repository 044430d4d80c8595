"""Placeholder factory -- mirror of ``kgcn/default_model.py`` (``DefaultModel.get_placeholders``,
lines 7-41): declares the feed signature (``adjs`` = ``batch_size x adj_channel_num`` sparse inputs,
``features [B,N,F]``, ``labels``, ``mask``, ``enabled_node_nums`` ...).  Placeholders here are inert
descriptors; ``kgcn_b200.compat.ModelRunner`` replaces them by the step's tensors."""


def _ph(dtype, shape=None, name=None, sparse=False):
    from .compat.facade import Placeholder
    return Placeholder(dtype, shape, name, sparse)


class DefaultModel:
    def get_placeholders(self, info, config, batch_size, placeholder_names, **kwargs):
        C, N = info.adj_channel_num, info.graph_node_num
        g = lambda k, d=None: getattr(info, k, d) if not isinstance(info, dict) else info.get(k, d)
        label_dim = g("label_dim", 0)
        placeholders = {
            "adjs": [[_ph("float32", name="adj_%d_%d" % (a, b), sparse=True) for a in range(C)] for b in range(batch_size)],
            "nodes": _ph("int32", (batch_size, N), "node"),
            "node_label": _ph("float32", (batch_size, N, label_dim), "node_label"),
            "mask_node_label": _ph("float32", (batch_size, N, label_dim), "node_mask_label"),
            "labels": _ph("float32", (batch_size, label_dim), "label"),
            "mask": _ph("float32", (batch_size,), "mask"),
            "mask_label": _ph("float32", (batch_size, label_dim), "mask_label"),
            "mask_node": _ph("float32", (batch_size, N), "mask_node"),
            "dropout_rate": _ph("float32", name="dropout_rate"),
            "enabled_node_nums": _ph("int32", (batch_size,), "enabled_node_nums"),
            "is_train": _ph("bool", name="is_train"),
        }
        placeholders["features"] = _ph("float32", (batch_size, N, g("feature_dim", 0)), "feature") if g("feature_enabled", True) else None
        # the remaining names of the reference's table (default_model.py:22-39): inputs of the multimodal / sequence /
        # link-prediction model families.  Declared so that any model file's key list resolves; the graph path feeds none.
        seq_len = g("sequence_max_length", 0)
        placeholders.update({
            "sequences": _ph("int32", (batch_size, seq_len), "sequences"),
            "sequences_vec": _ph("float32", (batch_size, seq_len, g("sequences_vec_dim", 0)), "sequences_vec"),
            "sequences_len": _ph("int32", (batch_size, 2), "sequences_len"),
            "preference_label_list": _ph("int64", (batch_size, None, 6), "preference_label_list"),
            "label_list": _ph("int64", (batch_size, None, 2), "label_list"),
            "embedded_layer": _ph("float32", (batch_size, seq_len, (config or {}).get("embedding_dim")), "embedded_layer"),
        })
        modal_dims = g("vector_modal_dim", None) or []
        for name, j in (g("vector_modal_name", None) or {}).items():
            placeholders[name] = _ph("float32", (batch_size, modal_dims[j]), name)
        self.placeholders = {name: placeholders[name] for name in placeholder_names}
        return self.placeholders
