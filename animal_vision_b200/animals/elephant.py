"""Drop-in for reference animals/elephant.py."""
from .mammals import Elephant  # noqa: F401
