"""TEST-ONLY shim of SI_Toolkit.load_and_normalize.load_yaml."""
import yaml


def load_yaml(path, mode="r"):
    with open(path, mode) as f:
        return yaml.load(f, Loader=yaml.FullLoader)
