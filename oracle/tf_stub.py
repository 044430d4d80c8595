"""Minimal ``tensorflow`` stand-in so the reference's *ingest* code imports unchanged.

TEST INFRASTRUCTURE ONLY.  Nothing under ``kgcn_b200/`` imports this module.

``/root/reference/kgcn/data_util.py`` and ``/root/reference/kgcn/feed.py`` only touch two
TensorFlow symbols at import/run time on the paths we exercise: ``tf.__version__``
(feed.py:1-5 / data_util has none) and ``tf.SparseTensorValue`` (feed.py:122,126).  Installing
this stub lets ``oracle/make_golden.py`` run the reference's own numpy code to produce the
golden ingest vectors committed under ``tests/golden/``.
"""
import collections
import sys
import types

SparseTensorValue = collections.namedtuple("SparseTensorValue", ["indices", "values", "dense_shape"])


def install():
    """Put the stub on ``sys.modules['tensorflow']`` (idempotent) and return it."""
    mod = sys.modules.get("tensorflow")
    if mod is not None and getattr(mod, "_kgcn_b200_stub", False):
        return mod
    mod = types.ModuleType("tensorflow")
    mod.__version__ = "1.15.0"
    mod.SparseTensorValue = SparseTensorValue
    mod._kgcn_b200_stub = True
    sys.modules["tensorflow"] = mod
    return mod
