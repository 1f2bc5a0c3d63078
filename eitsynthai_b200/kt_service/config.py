"""Reads ai_fsi_config.toml (the reference reads it only in scripts/ai_fsi_logger.py:5 and never
uses ``[ai_settings] device``; here that key selects the GPU)."""
from __future__ import annotations

import os

_HERE = os.path.dirname(os.path.abspath(__file__))


def load() -> dict:
    path = os.environ.get("AI_FSI_CONFIG", os.path.join(_HERE, "ai_fsi_config.toml"))
    try:
        import tomllib
        with open(path, "rb") as f:
            return tomllib.load(f)
    except Exception:
        return {"main_settings": {"service_version": "1.0", "save_log_path": ["ai_logs"]},
                "ai_settings": {"device": "", "weights_ribs": "/", "weights_segmentation": "/"}}


def device() -> str:
    d = load().get("ai_settings", {}).get("device", "")
    return d if d else "cuda:0"
