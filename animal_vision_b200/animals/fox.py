"""Drop-in for reference animals/fox.py."""
from .mammals import Fox  # noqa: F401
