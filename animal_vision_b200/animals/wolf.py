"""Drop-in for reference animals/wolf.py."""
from .mammals import Wolf  # noqa: F401
