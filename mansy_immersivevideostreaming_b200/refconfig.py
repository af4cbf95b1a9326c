"""Loader for the reference's ``config.yml`` without its ``munch`` dependency.

Mirrors ``get_config_from_yml`` (bitrate_selection/utils/common.py:13-37): the YAML document becomes an
attribute-access dict and the dataset / results / models sub-directories are prefixed with their base
directories.  The reference's own ``Munch`` config objects are accepted everywhere this one is.
"""
from __future__ import annotations

from typing import Optional

DEFAULT_CONFIG_YML_PATH = "../config.yml"      # utils/common.py:10 (scripts run from bitrate_selection/)


class AttrDict(dict):
    """``dict`` with attribute access (what the reference uses ``munch.Munch`` for)."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as exc:            # keep hasattr()/getattr(default) semantics
            raise AttributeError(key) from exc

    def __setattr__(self, key, value):
        self[key] = value


def load_config_yml(config_yml_path: Optional[str] = None) -> AttrDict:
    import yaml
    path = config_yml_path or DEFAULT_CONFIG_YML_PATH
    with open(path, "r", encoding="utf8") as fh:
        config = AttrDict(yaml.load(fh, Loader=yaml.SafeLoader))
    for key in ("raw_datasets_dir", "raw_network_datasets_dir", "viewport_datasets_dir", "video_datasets_dir",
                "network_datasets_dir"):
        for name in config[key]:
            config[key][name] = config.datasets_base_dir + config[key][name]
    config.vp_results_dir = config.results_base_dir + config.vp_results_dir
    config.bs_results_dir = config.results_base_dir + config.bs_results_dir
    config.vp_models_dir = config.models_base_dir + config.vp_models_dir
    config.bs_models_dir = config.models_base_dir + config.bs_models_dir
    return config
