"""``flashinfer.version`` (reference flashinfer/version.py): the package version string."""
__version__ = "0.1.0"
__git_version__ = "unknown"
