"""Drop-in for reference animals/lion.py."""
from .mammals import Lion  # noqa: F401
