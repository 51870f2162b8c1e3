"""Drop-in for reference animals/squirrel.py."""
from .mammals import Squirrel  # noqa: F401
