"""Drop-in for reference animals/horse.py."""
from .mammals import Horse  # noqa: F401
