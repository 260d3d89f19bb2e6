"""Import path of the reference package, kept so ``run.py`` and callers work unchanged.

The implementation lives in ``shoeprint-image-retrieval_b200/`` at the repository root (a name
Python cannot import directly); it is appended to this package's ``__path__`` so that
``src.shoeprint_image_retrieval.similarity`` etc. resolve to the B200-native modules.
"""

from pathlib import Path as _Path

_IMPL = _Path(__file__).resolve().parents[2] / "shoeprint-image-retrieval_b200"
if not _IMPL.is_dir():  # pragma: no cover - broken checkout
    raise ImportError(f"implementation directory {_IMPL} is missing")
__path__.append(str(_IMPL))
