"""languagegroundedsemseg_b200 — B200-native sparse-voxel convolution engine behind the MinkowskiEngine call sites
of RozDavid/LanguageGroundedSemseg (hot path only)."""
from . import _lib  # noqa: F401
from . import minkowski  # noqa: F401


def install_as_minkowski(name="MinkowskiEngine"):
    """Make `import MinkowskiEngine` resolve to the lgs_b200 facade (the reference's models/ then run unchanged)."""
    return minkowski.install(name)


def load_reference_checkpoint(model, checkpoint, lenient=True, map_location="cpu"):
    """Load a checkpoint written by the reference (lib/utils.py:48-70 / Lightning) into an engine network."""
    from .checkpoint import load_reference_checkpoint as _load
    return _load(model, checkpoint, lenient=lenient, map_location=map_location)


__all__ = ["minkowski", "install_as_minkowski", "load_reference_checkpoint"]
