"""Label statistics (drop-in for ``upcp.analysis.analysis_tools.get_label_stats``,
reference src/upcp/analysis/analysis_tools.py:8-18).  The histogram is computed on
the device for device-resident labels; it is also the payload of the end-of-batch
NCCL gather (see ``batch.py``)."""
import numpy as np
import torch

from .. import device as dev
from ..labels import Labels


def label_histogram(labels):
    """int64[256] histogram of label values (numpy)."""
    if isinstance(labels, torch.Tensor):
        return dev.label_hist(labels).cpu().numpy()
    d_labels = dev.upload_labels(labels, dev.current_device())
    return dev.label_hist(d_labels).cpu().numpy()


def format_label_stats(hist, n=None):
    """Render a histogram the way the reference prints it."""
    hist = np.asarray(hist)
    if n is None:
        n = int(hist.sum())
    stats = f'Total: {n:25} points\n'
    for label in np.nonzero(hist)[0]:
        cnt = int(hist[label])
        name = Labels.STR_DICT.get(int(label), 'Unknown')
        perc = (float(cnt) / n) * 100
        stats += f'Class {label:2}, {name:14} ' + f'{cnt:7} points ({perc:4.1f} %)\n'
    return stats


def get_label_stats(labels):
    """Returns a string describing statistics based on labels."""
    n = int(labels.numel()) if isinstance(labels, torch.Tensor) else len(labels)
    return format_label_stats(label_histogram(labels), n)
