"""Plugin contract of the labelling pipeline (drop-in for
``upcp.abstract_processor.AbstractProcessor``, reference
src/upcp/abstract_processor.py:8-47).

Every processor exposes
  * ``get_labels(points, labels, mask, tilecode) -> labels`` -- exactly the
    reference contract: ``labels`` is mutated in place and returned;
  * ``get_label_mask(points, labels, mask, tilecode) -> bool[N]`` -- the uniform
    mask-returning form named by the north-star (the reference only has it on
    LabelConnectedComp, label_connected_comp.py:137-181);
  * ``_label_mask_device(tile, d_labels, d_mask, tilecode) -> uint8 cuda tensor`` --
    the device-resident form the Pipeline chains without leaving HBM.
"""
import threading
from abc import ABC, abstractmethod

import numpy as np

from . import device as dev


class PerThread:
    """Diagnostic attribute with one value per host thread.

    ``Pipeline.process_clouds`` / ``process_folder(workers=W)`` drive ONE processor object from W threads, each on
    its own tile; what a call leaves behind for inspection (``last_timings``, ``last_stats``, ``last_objects``,
    ``octree_level``) belongs to the thread that made the call, so that a worker never reads -- or, in
    ``CarFuser``, clips with -- another tile's value.  A thread that has not set the attribute reads the default."""

    def __init__(self, factory):
        self.factory = factory

    def __set_name__(self, owner, name):
        self.name = name

    def _store(self, obj):
        store = obj.__dict__.get('_per_thread')
        if store is None:
            store = obj.__dict__.setdefault('_per_thread', threading.local())
        return store

    def __get__(self, obj, owner=None):
        if obj is None:
            return self
        store = self._store(obj)
        if not hasattr(store, self.name):
            setattr(store, self.name, self.factory())
        return getattr(store, self.name)

    def __set__(self, obj, value):
        setattr(self._store(obj), self.name, value)


def drop_per_thread_state(obj):
    """``__getstate__`` helper: the per-thread store is a ``threading.local`` and does not pickle; a copy of the
    object (``copy.deepcopy``, ``pickle`` for a process pool) starts with empty diagnostics."""
    state = obj.__dict__.copy()
    state.pop('_per_thread', None)
    return state


class AbstractProcessor(ABC):
    """Abstract base class for automatic labelling of point clouds.

    Parameters
    ----------
    label : int
        Class label to use for this processor.
    """

    def __init__(self, label):
        self.label = label
        super().__init__()

    def __getstate__(self):
        return drop_per_thread_state(self)

    # ---- device-resident core, implemented by every drop-in -----------------
    @abstractmethod
    def _label_mask_device(self, tile, d_labels, d_mask, tilecode):
        """Return a uint8 cuda tensor [N] (1 = point receives ``self.label``) or
        ``None`` when the processor has nothing to do for this tile."""

    def _apply_device(self, tile, d_labels, d_mask, tilecode):
        """labels[label_mask] = label, on the device."""
        d_label_mask = self._label_mask_device(tile, d_labels, d_mask, tilecode)
        if d_label_mask is not None:
            dev.labels_assign(d_labels, d_label_mask, self.label)
        return d_labels

    # ---- reference-facing numpy API ------------------------------------------
    def get_label_mask(self, points, labels, mask=None, tilecode=None):
        """Boolean mask of the points this processor labels (numpy in, numpy out)."""
        tile = dev.get_tile(points)
        d_labels = dev.upload_labels(labels, tile.device)
        d_mask = dev.upload_mask(mask, tile.n, tile.device)
        d_label_mask = self._label_mask_device(tile, d_labels, d_mask, tilecode)
        if d_label_mask is None:
            return np.zeros((tile.n,), dtype=bool)
        return dev.download_mask(d_label_mask)

    def get_labels(self, points, labels, mask, tilecode):
        """Returns the labels for the given pointcloud (updated in place)."""
        tile = dev.get_tile(points)
        d_labels = dev.upload_labels(labels, tile.device)
        d_mask = dev.upload_mask(mask, tile.n, tile.device)
        d_label_mask = self._label_mask_device(tile, d_labels, d_mask, tilecode)
        if d_label_mask is None:
            return labels
        dev.labels_assign(d_labels, d_label_mask, self.label)
        return dev.download_labels(d_labels, labels)

    def get_label(self):
        """Returns the label of this AbstractProcessor."""
        return self.label
