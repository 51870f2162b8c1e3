"""Drop-in for reference animals/rat.py."""
from .mammals import Rat  # noqa: F401
