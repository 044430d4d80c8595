"""A model definition written against the REFERENCE's API (``import tensorflow as tf``,
``import kgcn.layers``, ``DefaultModel``) -- the model-module protocol of gcn.py:135-151 /
kgcn/core.py:156-157.  It is a test fixture for ``kgcn_b200.compat``: the file contains no
kgcn_b200 import and runs unchanged on the façade."""
import tensorflow as tf

if tf.__version__.split(".")[0] == "2":
    import tensorflow.compat.v1 as tf
    tf.disable_v2_behavior()
    import tensorflow.keras as K
else:
    import tensorflow.contrib.keras as K

import kgcn.layers
from kgcn.default_model import DefaultModel


class Net(DefaultModel):
    def build_placeholders(self, info, config, batch_size, **kwargs):
        keys = ["adjs", "labels", "mask", "dropout_rate", "enabled_node_nums", "is_train", "features"]
        return self.get_placeholders(info, config, batch_size, keys, **kwargs)

    def build_model(self, placeholders, info, config, batch_size, **kwargs):
        adjs, x = placeholders["adjs"], placeholders["features"]
        labels, mask = placeholders["labels"], placeholders["mask"]
        h = tf.sigmoid(kgcn.layers.GraphConv(16, info.adj_channel_num)(x, adj=adjs))
        h = kgcn.layers.GraphConv(12, info.adj_channel_num)(h, adj=adjs)
        h = kgcn.layers.GraphBatchNormalization()(h, max_node_num=info.graph_node_num,
                                                  enabled_node_nums=placeholders["enabled_node_nums"])
        h = tf.nn.relu(h)
        h = K.layers.Dropout(placeholders["dropout_rate"])(h)
        h = tf.sigmoid(kgcn.layers.GraphDense(8)(h))
        g = kgcn.layers.GraphGather()(h)
        logits = K.layers.Dense(info.label_dim)(g)
        prediction = tf.nn.softmax(logits, name="output")
        cost = mask * tf.nn.softmax_cross_entropy_with_logits_v2(labels=labels, logits=logits)
        cost_opt = tf.reduce_mean(cost)
        cost_sum = tf.reduce_sum(cost)
        correct = mask * tf.cast(tf.equal(tf.argmax(prediction, 1), tf.argmax(labels, 1)), tf.float32)
        self.out = logits
        return self, prediction, cost_opt, cost_sum, {"correct_count": tf.reduce_sum(correct)}
