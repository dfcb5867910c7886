"""Config loading — host mirror of PKG/helpers.py:4-24 (`load_config`).

`validate_config` / `print_config` of the reference are dead code (never called; the
former requires a section the YAML lacks) and are not mirrored.
"""
from __future__ import annotations

import os

import yaml

PACKAGE_PARAMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mpc_parameters.yaml")


def load_config(config_file):
    """YAML file -> dict.  Like the reference the path is taken as given (relative to the CWD);
    unlike it, a missing file falls back to the packaged copy of mpc_parameters.yaml."""
    path = config_file
    if not os.path.exists(path) and os.path.basename(str(config_file)) == "mpc_parameters.yaml":
        path = PACKAGE_PARAMS
    with open(path, "r", encoding="utf-8") as f:
        return yaml.safe_load(f)
