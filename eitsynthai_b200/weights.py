"""Checkpoint loader for the three networks of the service.

The reference loads its weights with ``YOLO(model_path, task='segment')``
(kt_service/ai_tools/ai_tools.py:69-71; paths from kt_service_config.py:1-3).  An ultralytics ``.pt``
file is a pickle of ``{'model': SegmentationModel, 'ema': ..., ...}`` whose tensors live in the module
tree under the names ``model.<layer>.<sub-module>...`` of ``yolo11-seg.yaml``.  ultralytics is not a
dependency here, so the pickle is read with stand-in classes for everything under ``ultralytics.*``
(only the ``_parameters`` / ``_buffers`` / ``_modules`` dictionaries of each module are kept), the
tensors are renamed onto ``YOLO11sSeg`` (``model.N`` -> ``lN``, ``model.23`` -> ``head``), the real
BatchNorm statistics are loaded and then folded by ``fuse()``.  A plain ``state_dict`` saved with the
same key names is accepted too.
"""
from __future__ import annotations

import io
import pickle
import re
import types

import torch

from .yolo_seg import YOLO11sSeg, finalize_model


class CheckpointError(RuntimeError):
    pass


# ------------------------------------------------------------------------------------ reading
class _Stub:
    """Stands in for any class the pickle names that is not importable (ultralytics modules, loss objects...)."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        elif isinstance(state, tuple) and len(state) == 2 and isinstance(state[0], dict):   # (dict, slots)
            self.__dict__.update(state[0])
            if isinstance(state[1], dict):
                self.__dict__.update(state[1])

    def __call__(self, *a, **k):                                # reduce-style callables
        return _Stub()


class _StubUnpickler(pickle.Unpickler):
    _ALLOWED_PREFIX = ("torch", "collections", "numpy", "builtins", "_codecs", "copyreg", "pathlib", "datetime")

    def find_class(self, module, name):
        if module.split(".")[0] in self._ALLOWED_PREFIX:
            try:
                return super().find_class(module, name)
            except (ImportError, AttributeError):
                pass
        return type(name, (_Stub,), {"__module__": module})


_stub_pickle = types.ModuleType("eitb_stub_pickle")
_stub_pickle.Unpickler = _StubUnpickler
_stub_pickle.load = lambda f, **kw: _StubUnpickler(f, **kw).load()
_stub_pickle.loads = lambda b, **kw: _StubUnpickler(io.BytesIO(b), **kw).load()
_stub_pickle.__name__ = "pickle"
for _n in ("PickleError", "PicklingError", "UnpicklingError", "HIGHEST_PROTOCOL", "DEFAULT_PROTOCOL", "dump", "dumps", "Pickler"):
    setattr(_stub_pickle, _n, getattr(pickle, _n))


def _walk(obj, prefix: str, out: dict):
    """Collect the tensors of a (stand-in or real) module tree under state_dict-style names."""
    d = getattr(obj, "__dict__", {})
    for kind in ("_parameters", "_buffers"):
        for k, v in (d.get(kind) or {}).items():
            if isinstance(v, torch.Tensor):
                out[prefix + k] = v.detach()
    for k, child in (d.get("_modules") or {}).items():
        if child is not None:
            _walk(child, f"{prefix}{k}.", out)


def read_state(path: str) -> dict:
    """{ultralytics-style key: tensor} from an ultralytics ``.pt`` checkpoint or a saved state_dict."""
    try:
        obj = torch.load(path, map_location="cpu", weights_only=False, pickle_module=_stub_pickle)
    except Exception as e:
        raise CheckpointError(f"cannot read checkpoint {path}: {e}") from e
    if isinstance(obj, dict) and obj and all(isinstance(v, torch.Tensor) for v in obj.values()):
        return {k: v.detach() for k, v in obj.items()}
    if isinstance(obj, dict):
        root = obj.get("ema") or obj.get("model")
        if isinstance(root, dict):                             # {'model': state_dict}
            return {k: v.detach() for k, v in root.items() if isinstance(v, torch.Tensor)}
    else:
        root = obj
    if root is None:
        raise CheckpointError(f"{path}: no 'model' / 'ema' entry")
    out: dict = {}
    _walk(root, "", out)
    if not out:
        raise CheckpointError(f"{path}: no tensors found in the module tree")
    return out


# ------------------------------------------------------------------------------------ renaming
_LAYER = re.compile(r"^(?:model\.)?model\.(\d+)\.(.*)$")


def map_keys(state: dict) -> dict:
    """ultralytics names -> ``YOLO11sSeg`` names; drops DFL's constant conv and BatchNorm counters."""
    out = {}
    for k, v in state.items():
        m = _LAYER.match(k)
        if not m:
            continue
        idx, rest = int(m.group(1)), m.group(2)
        if rest.endswith("num_batches_tracked") or rest.startswith("dfl."):
            continue
        out[("head." if idx == 23 else f"l{idx}.") + rest] = v
    return out


def load_model(path: str, device, dtype=torch.float16, nc: int | None = None) -> YOLO11sSeg:
    """Build ``YOLO11sSeg`` from a checkpoint: real weights and BatchNorm statistics, then fused like the
    random-init models of ``yolo_seg.build_model``."""
    mapped = map_keys(read_state(path))
    last = [k for k in mapped if re.match(r"^head\.cv3\.0\.2\.weight$", k)]
    if not last:
        raise CheckpointError(f"{path}: not a YOLO11-seg checkpoint (no model.23.cv3.0.2.weight)")
    nc_ckpt = int(mapped[last[0]].shape[0])
    if nc is not None and nc != nc_ckpt:
        raise CheckpointError(f"{path}: checkpoint has {nc_ckpt} classes, expected {nc}")
    model = YOLO11sSeg(nc_ckpt).eval()
    own = model.state_dict()
    missing = [k for k in own if k not in mapped and not k.endswith("num_batches_tracked")]
    if missing:
        raise CheckpointError(f"{path}: {len(missing)} tensors missing, e.g. {missing[:3]}")
    bad = [k for k in own if k in mapped and tuple(mapped[k].shape) != tuple(own[k].shape)]
    if bad:
        raise CheckpointError(f"{path}: shape mismatch (not the 's' scale?), e.g. {bad[0]}: "
                              f"{tuple(mapped[bad[0]].shape)} vs {tuple(own[bad[0]].shape)}")
    model.load_state_dict({k: mapped[k].float() for k in own if k in mapped}, strict=False)
    return finalize_model(model, device, dtype, fuse=True)
