"""TOML configuration files -> plain dictionaries (schema of config/model-unet.toml and config/dataset-*.toml)."""

try:
    import toml as _toml

    def load_config(path):
        return _toml.load(path)

    def save_config(attrs, path):
        with open(path, "w") as fp:
            _toml.dump(attrs, fp)

except ImportError:  # pragma: no cover - Python 3.11+ ships a reader
    import tomllib

    def load_config(path):
        with open(path, "rb") as fp:
            return tomllib.load(fp)

    def save_config(attrs, path):
        raise RuntimeError("writing TOML needs the `toml` package")
