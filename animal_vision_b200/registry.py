"""Species registry in the shape the reference's glue expects (utils.py:91-130 `animal_choices`:
a list of {"name": str, "value": Animal instance}; names as the reference spells them).

Covers the species this implementation runs on the GPU: the 20 non-UV mammals and HoneyBee.
The other 15 UV species of the reference are not implemented yet and are simply
absent -- there is no CPU fallback to stand in for them."""
from __future__ import annotations

from typing import Dict, List

from . import animals as A

# order and spelling of utils.py:91-112
_NAMES = ["Cat", "Dog", "Sheep", "Pig", "Goat", "Cow", "Horse", "Rabbit", "Panda", "Squirrel", "Elephant", "Lion",
          "Wolf", "Fox", "Bear", "Raccoon", "Deer", "Kangaroo", "Tiger", "Rat", "HoneyBee"]


def animal_classes() -> Dict[str, type]:
    return {n: getattr(A, n) for n in _NAMES}


def animal_choices() -> List[dict]:
    """Instances are created lazily by the caller's first `visualize` (constructors touch no GPU)."""
    return [{"name": n, "value": cls()} for n, cls in animal_classes().items()]
