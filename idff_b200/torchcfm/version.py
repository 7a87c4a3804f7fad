"""Version of the torchcfm surface mirrored here (the reference's torchcfm/version.py reports the same string)."""
__version__ = ".".join(str(v) for v in (1, 0, 6))
