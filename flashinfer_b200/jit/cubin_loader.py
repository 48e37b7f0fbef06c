"""Artifact hooks of the reference (flashinfer/jit/cubin_loader.py).  That module downloads NVIDIA-built cubins (trtllm-gen FMHA /
GEMM / MoE) and registers a loader callback with the C++ runtime; this library has no pre-built binaries - every kernel is
compiled from ``csrc/`` - so the hooks answer "nothing to fetch" instead of touching the network."""
import hashlib
import os

FLASHINFER_CUBINS_REPOSITORY = os.environ.get("FLASHINFER_CUBINS_REPOSITORY", "")
dll_cubin_handlers: dict = {}


def get_meta_hash(checksums_bytes: bytes, target_file: str = "flashinferMetaInfo.h") -> str:
    """sha256 listed for ``target_file`` in a ``checksums.txt`` blob ("<sha256>  <path>" lines)."""
    for line in checksums_bytes.decode("utf-8", "replace").splitlines():
        parts = line.split()
        if len(parts) == 2 and parts[1].endswith(target_file):
            return parts[0]
    raise ValueError(f"{target_file} is not listed")


def verify_cubin(cubin_path: str, expected_sha256: str) -> bool:
    with open(cubin_path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest() == expected_sha256


def load_cubin(cubin_path: str, sha256: str) -> bytes:
    """Read a local file when it exists and matches ``sha256`` (empty bytes otherwise)."""
    if os.path.exists(cubin_path) and verify_cubin(cubin_path, sha256):
        with open(cubin_path, "rb") as f:
            return f.read()
    return b""


def get_artifact(file_name: str, sha256: str, session=None) -> bytes:
    """Nothing is downloaded: only a file already present under the in-tree library directory is returned."""
    from .env import FLASHINFER_CUBIN_DIR

    return load_cubin(str(FLASHINFER_CUBIN_DIR / file_name), sha256)


get_cubin = get_artifact


def setup_cubin_loader(dll_path: str = "") -> None:
    """The reference registers a ctypes callback that feeds cubins to its runtime; native modules here embed their SASS."""
    return None
