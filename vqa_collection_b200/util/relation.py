"""Drop-in for util/relation.py: relation graphs computed by the device kernel.

``relation_graph(bbox, w, h)`` keeps the reference signature and return type
(float64 [K,K], relation.py:65-80); ``relation_graph_batch`` is the batched form the
data pipeline should call (one launch for the whole batch)."""
import numpy as np

from .. import ops


def relation_graph_batch(bbox, w, h):
    """bbox [B,K,4] (numpy or CPU tensor) → uint8 numpy [B,K,K] (H2D + kernel + D2H)"""
    bbox = np.ascontiguousarray(np.asarray(bbox), dtype=np.float32)
    return ops.relation_labels_host(bbox, w, h)


def relation_graph(bbox, w, h, relation=None):
    if relation is not None:
        raise NotImplementedError("only the spatial relation is built (semantic_relation is a stub in the "
                                  "reference too, relation.py:48-62)")
    bbox = np.asarray(bbox)
    return relation_graph_batch(bbox[None], w, h)[0].astype(np.float64)


def spatial_relation(a, b, w, h):
    """label pair of two boxes (relation.py:3-45) through the same kernel"""
    lab = relation_graph_batch(np.stack([np.asarray(a), np.asarray(b)])[None], w, h)[0]
    return int(lab[0, 1]), int(lab[1, 0])
