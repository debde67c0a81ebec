"""Minimal reader for a workflow directory's config.json (the parts the SBS stage uses).

The reference's helper/config_manager.py (load_config :267, get_path :342, the `stereo` schema
:43-54) stays the owner of the format; this module reads the same file with the same rules for the
keys this stage needs, so that sbs_generator.py runs both inside the reference tree (where
`helper.config_manager` is importable and preferred) and standalone.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict

STEREO_KEYS = ('max_disparity', 'convergence', 'super_sampling', 'edge_softness', 'artifact_smoothing', 'depth_gamma', 'sharpen')
DIR_KEYS = ('frames', 'depth_maps', 'sbs')


class ConfigError(Exception):
    """Raised when config.json is missing or fails validation (reference: config_manager.py:78)."""


def load_config(workflow_path) -> Dict:
    f = Path(workflow_path) / 'config.json'
    if not f.exists():
        raise ConfigError(f'Config file not found: {f}')
    try:
        cfg = json.loads(f.read_text(encoding='utf-8'))
    except json.JSONDecodeError as e:
        raise ConfigError(f'Invalid JSON in config file: {e}')
    if not isinstance(cfg.get('directories'), dict):
        raise ConfigError("Missing required key: 'directories'")
    for k in DIR_KEYS:
        if not isinstance(cfg['directories'].get(k), str):
            raise ConfigError(f"Missing or invalid 'directories.{k}' (expected str)")
    if not isinstance(cfg.get('stereo'), dict):
        raise ConfigError("Missing required key: 'stereo'")
    for k in STEREO_KEYS:
        v = cfg['stereo'].get(k)
        # ints are accepted for floats (config_manager.py:114); bools are not numbers here
        if isinstance(v, bool) or not isinstance(v, (int, float)):
            raise ConfigError(f"Missing or invalid 'stereo.{k}' (expected float)")
    return cfg


def get_path(workflow_path, config: Dict, key: str) -> Path:
    if key not in config['directories']:
        raise KeyError(f'Unknown directory key: {key}')
    return Path(workflow_path) / config['directories'][key]


def find_frame_pairs(frames_dir: Path, depth_dir: Path):
    """Matching (frame_path, depth_path, frame_num) triples; .tif depth preferred over .png
    (reference: sbs_generator.py:71-116).  Returns (pairs, missing_count, first_missing, last_missing)."""
    pairs, missing, first, last = [], 0, None, None
    for frame_path in sorted(Path(frames_dir).glob('frame_*.png')):
        num = frame_path.stem.replace('frame_', '')
        dp = Path(depth_dir) / f'depth_frame_{num}.tif'
        if not dp.exists():
            dp = Path(depth_dir) / f'depth_frame_{num}.png'
            if not dp.exists():
                first = first if first is not None else num
                last = num
                missing += 1
                continue
        pairs.append((frame_path, dp, num))
    return pairs, missing, first, last
