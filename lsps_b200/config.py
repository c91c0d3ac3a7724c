"""Config surface of the reference: every key under YAML `train:` becomes an attribute
(/root/reference/src/utils/net_config.py:9-20; key list in exps/nnyu.yaml)."""
import os

import yaml

EXPS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "exps")


class NetConfig(object):
    def __init__(self, config):
        with open(config, "r") as fh:
            doc = yaml.safe_load(fh)
        for k, v in doc["train"].items():
            setattr(self, k, v)


def load_hyperparameters(name="nnyu"):
    path = name if os.path.exists(name) else os.path.join(EXPS, name + ".yaml")
    return NetConfig(path).hyperparameters
