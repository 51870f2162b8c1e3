"""Drop-in for reference animals/pig.py."""
from .mammals import Pig  # noqa: F401
