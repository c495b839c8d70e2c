"""Import stub of omegaconf (absent; model_training/utils.py:5)."""


class DictConfig(dict):
    pass


class OmegaConf:
    @staticmethod
    def to_yaml(cfg, resolve=True):
        import yaml
        return yaml.dump(dict(cfg))
