"""Lip-vertex error as the reference computes it (TEST INFRASTRUCTURE): metric/metric.py:120-123,136 —
per frame, the maximum over the lip vertices of the squared L2 distance; then the mean over frames."""
import numpy as np


def lip_vertex_error(gt: np.ndarray, pred: np.ndarray, lip_idx: np.ndarray) -> float:
    """gt, pred: (frames, V*3)."""
    g = gt.reshape(gt.shape[0], -1, 3)[:, lip_idx]
    p = pred.reshape(pred.shape[0], -1, 3)[:, lip_idx]
    return float(np.mean(np.max(np.sum(np.square(g - p), axis=2), axis=1)))
