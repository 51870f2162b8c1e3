"""Drop-in for reference animals/goat.py."""
from .mammals import Goat  # noqa: F401
