"""languagegroundedsemseg_b200 — B200-native sparse-voxel convolution engine behind the MinkowskiEngine call sites
of RozDavid/LanguageGroundedSemseg (hot path only)."""
from . import _lib  # noqa: F401
from . import minkowski  # noqa: F401


def install_as_minkowski(name="MinkowskiEngine"):
    """Make `import MinkowskiEngine` resolve to the lgs_b200 facade (the reference's models/ then run unchanged)."""
    return minkowski.install(name)


__all__ = ["minkowski", "install_as_minkowski"]
