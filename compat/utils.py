"""reference path: utils.py (utils.py:1-13)"""
import os
from typing import Any, Dict

import yaml


def get_relative_path(x: str, rel_to: str) -> str:
    return os.path.join(os.path.dirname(rel_to), x)


def load_yaml(x: str) -> Dict[str, Any]:
    with open(x) as fd:
        return yaml.load(fd, yaml.FullLoader)
