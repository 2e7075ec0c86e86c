"""pesq stand-in: importable, unusable (quality metrics / training data are not on the inference path)."""


def __getattr__(name):
    def _missing(*args, **kwargs):
        raise NotImplementedError("pesq." + name + ": not available in this image")
    _missing.__name__ = name
    return _missing
