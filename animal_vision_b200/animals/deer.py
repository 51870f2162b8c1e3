"""Drop-in for reference animals/deer.py."""
from .mammals import Deer  # noqa: F401
