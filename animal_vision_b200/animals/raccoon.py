"""Drop-in for reference animals/raccoon.py."""
from .mammals import Raccoon  # noqa: F401
