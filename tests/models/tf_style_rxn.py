"""The architecture the reference's shipped checkpoint (model/reaction/model.best.ckpt.*) was trained with --
three legacy GraphConv(128) + legacy GraphBatchNormalization + relu blocks, GraphDense(128), GraphGather,
Dense(label_dim) under ``tf.variable_scope("rollout")`` -- written against the REFERENCE's API only
(``import tensorflow as tf``, ``import kgcn.legacy.layers``, ``DefaultModel``).  Test fixture for
``kgcn_b200.compat``: no kgcn_b200 import; the variable names TensorFlow would create for this file are the
checkpoint's (``rollout/graph_conv_1/kernel0``, ``rollout/batch_normalization_2/gamma``, ``rollout/dense/bias`` ...)."""
import tensorflow as tf

if tf.__version__.split(".")[0] == "2":
    import tensorflow.compat.v1 as tf
    tf.disable_v2_behavior()
    import tensorflow.keras as K
else:
    import tensorflow.contrib.keras as K

import kgcn.legacy.layers as L
from kgcn.default_model import DefaultModel


class Rollout(DefaultModel):
    def build_placeholders(self, info, config, batch_size, **kwargs):
        keys = ["adjs", "labels", "mask", "enabled_node_nums", "is_train", "features"]
        return self.get_placeholders(info, config, batch_size, keys, **kwargs)

    def build_model(self, placeholders, info, config, batch_size, **kwargs):
        adjs, labels, mask = placeholders["adjs"], placeholders["labels"], placeholders["mask"]
        h = placeholders["features"]
        with tf.variable_scope("rollout"):
            for _ in range(3):
                h = L.GraphConv(128, info.adj_channel_num)(h, adj=adjs)
                h = L.GraphBatchNormalization()(h, max_node_num=info.graph_node_num,
                                                enabled_node_nums=placeholders["enabled_node_nums"])
                h = tf.nn.relu(h)
            h = tf.nn.relu(L.GraphDense(128)(h))
            self.gathered = L.GraphGather()(h)
            logits = K.layers.Dense(info.label_dim)(self.gathered)
            prediction = tf.nn.softmax(logits, name="output")
            cost = mask * tf.nn.softmax_cross_entropy_with_logits_v2(labels=labels, logits=logits)
            correct = mask * tf.cast(tf.equal(tf.argmax(prediction, 1), tf.argmax(labels, 1)), tf.float32)
        self.out = logits
        return self, prediction, tf.reduce_mean(cost), tf.reduce_sum(cost), {"correct_count": tf.reduce_sum(correct)}
