"""Drop-in for reference animals/tiger.py."""
from .mammals import Tiger  # noqa: F401
