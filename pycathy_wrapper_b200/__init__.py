"""cathy-b200: B200-native CATHY Richards-equation processor behind pyCATHY's run_processor boundary."""
from .project import CathyInputError, CathyProject, load_project  # noqa: F401

__version__ = "0.1.0"


def launcher_path() -> str:
    """Path of the `cathy` launcher script to drop into a pyCATHY project directory."""
    import os
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "cathy")
