"""Pre-compiled model tables (derived from the reference's MJCF by ``tools/compile_models.py``)."""

from pathlib import Path

from ..model import Model

_HERE = Path(__file__).resolve().parent
NAMES = ("franka_scene", "franka_scene_with_obstacles", "ur5e_scene", "one_dof_ball", "two_dof_ball")


def load(name: str) -> Model:
    """Load one of the bundled models (``NAMES``)."""
    if name not in NAMES:
        raise KeyError(f"unknown bundled model '{name}', choose from {NAMES}")
    return Model.load(_HERE / f"{name}.npz")
