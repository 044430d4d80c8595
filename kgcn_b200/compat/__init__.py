"""Compatibility façade: load kGCN model definitions (``example_model/*.py``) UNCHANGED.

Those files ``import tensorflow as tf`` / ``import kgcn.layers`` and describe a TF-1.x graph
(``build_placeholders`` / ``build_model``, the model-module protocol of gcn.py:135-151 and
kgcn/core.py:156-157).  :func:`install` puts a small ``tensorflow``-named module (eager, torch-backed)
and a ``kgcn`` alias package (-> ``kgcn_b200.layers`` / ``kgcn_b200.default_model``) on
``sys.modules``; :class:`ModelRunner` then runs ``build_model`` eagerly once per step with the
placeholders replaced by the step's tensors.  Layers created inside ``build_model`` obtain their
weights from a variable store keyed by TF-style names (``graph_conv_1/kernel0`` ...), so the second
and later steps reuse the parameters of the first -- what ``tf.variable_scope`` reuse does in the
reference.  This is NOT a TensorFlow runtime: no graph mode, no sessions, only the ~40 symbols the
in-scope model files touch.
"""
from .facade import ModelRunner, VariableStore, install, uninstall  # noqa: F401
