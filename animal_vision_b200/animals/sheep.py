"""Drop-in for reference animals/sheep.py."""
from .mammals import Sheep  # noqa: F401
