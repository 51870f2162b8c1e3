"""Drop-in for reference animals/bear.py."""
from .mammals import Bear  # noqa: F401
