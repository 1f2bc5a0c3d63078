"""Result objects with the attribute layout the reference reads from its third-party packages.

``ai_tools.py:121-123,153`` gets an ``ultralytics.engine.results.Results`` from ``model(...)[0]`` and wraps the rib
detections in ``sv.Detections.from_ultralytics``; the functions of ``utils.py`` then read

    results.masks.data (n, S, S) / results.boxes.cls / .conf / .xyxy / results.orig_shape      (utils.py:437-523)
    detections.xyxy / .confidence / .class_id / .mask                                         (utils.py:166-269)

Neither package is a dependency of this library; these classes carry the same fields (device tensors, ``.cpu()`` /
``.numpy()`` like the originals) so code written against the reference's objects runs unchanged on what K5 / K6 produce.
"""
from __future__ import annotations

import numpy as np
import torch


class _TensorBag:
    _fields = ()

    def cpu(self):
        return type(self)(*[getattr(self, f).cpu() if torch.is_tensor(getattr(self, f)) else getattr(self, f) for f in self._fields])

    def numpy(self):
        return type(self)(*[getattr(self, f).cpu().numpy() if torch.is_tensor(getattr(self, f)) else getattr(self, f) for f in self._fields])

    def __len__(self):
        return int(getattr(self, self._fields[0]).shape[0])


class Boxes(_TensorBag):
    """``results.boxes``: ``xyxy`` (n, 4), ``conf`` (n,), ``cls`` (n,) and ``data`` (n, 6) = xyxy | conf | cls."""
    _fields = ("xyxy", "conf", "cls")

    def __init__(self, xyxy, conf, cls):
        self.xyxy, self.conf, self.cls = xyxy, conf, cls

    @property
    def data(self):
        cat = torch.cat if torch.is_tensor(self.xyxy) else np.concatenate
        return cat([self.xyxy, self.conf[:, None], self.cls[:, None]], 1)


class Masks(_TensorBag):
    """``results.masks``: ``data`` (n, H, W) with 1 inside an instance's mask."""
    _fields = ("data",)

    def __init__(self, data):
        self.data = data


class Results:
    """One image's prediction: ``boxes``, ``masks`` (None without detections, like ultralytics), ``orig_shape``,
    ``names``."""

    def __init__(self, orig_shape, boxes: Boxes, masks: Masks | None, names: dict | None = None):
        self.orig_shape, self.boxes, self.masks = tuple(orig_shape), boxes, masks
        self.names = names or {}

    def __len__(self):
        return len(self.boxes)

    def cpu(self):
        return Results(self.orig_shape, self.boxes.cpu(), None if self.masks is None else self.masks.cpu(), self.names)


class Detections:
    """The fields of ``supervision.Detections`` the reference reads (numpy arrays, like supervision's)."""

    def __init__(self, xyxy, mask=None, confidence=None, class_id=None):
        self.xyxy = np.asarray(xyxy, np.float32).reshape(-1, 4)
        self.mask = mask
        self.confidence = np.zeros(len(self.xyxy), np.float32) if confidence is None else np.asarray(confidence, np.float32)
        self.class_id = np.zeros(len(self.xyxy), int) if class_id is None else np.asarray(class_id).astype(int)

    def __len__(self):
        return len(self.xyxy)

    @classmethod
    def from_ultralytics(cls, results: Results) -> "Detections":
        b = results.boxes.numpy()
        mask = None if results.masks is None else np.asarray(results.masks.numpy().data).astype(bool)
        return cls(b.xyxy, mask, b.conf, b.cls)

    @classmethod
    def empty(cls) -> "Detections":
        return cls(np.zeros((0, 4), np.float32))
