"""Species registry in the shape the reference's glue expects (utils.py:91-130 `animal_choices`:
a list of {"name": str, "value": Animal instance}; names, spelling and order as the reference has them).

All 36 entries of the reference run on the GPU: the 20 non-UV mammals, HoneyBee, the 14 UV species of SURVEY.md 8f-1
and MantisShrimp (8f-2).  There is no CPU fallback behind any of them."""
from __future__ import annotations

from typing import Dict, List

from . import animals as A

# (display name, class name): order and spelling of utils.py:91-130
_ENTRIES = [
    ("Cat", "Cat"), ("Dog", "Dog"), ("Sheep", "Sheep"), ("Pig", "Pig"), ("Goat", "Goat"), ("Cow", "Cow"), ("Horse", "Horse"),
    ("Rabbit", "Rabbit"), ("Panda", "Panda"), ("Squirrel", "Squirrel"), ("Elephant", "Elephant"), ("Lion", "Lion"), ("Wolf", "Wolf"),
    ("Fox", "Fox"), ("Bear", "Bear"), ("Raccoon", "Raccoon"), ("Deer", "Deer"), ("Kangaroo", "Kangaroo"), ("Tiger", "Tiger"),
    ("Rat", "Rat"),
    # UV based animals
    ("HoneyBee", "HoneyBee"), ("ReinDeer", "Reindeer"), ("RatUV", "RatUV"), ("GoldFish", "Goldfish"), ("DamselFish", "Damselfish"),
    ("Anableps (Four-eyed fish)", "Anableps"), ("Northern Anchovy Fish", "Anchovy"), ("Guppy Fish", "Guppy"),
    ("Morpho Butterfly", "Morpho"), ("Heliconius Butterfly", "Heliconius"), ("Pieris Butterfly", "Pieris"),
    # UV unique animals
    ("Mantis Shrimp", "MantisShrimp"), ("Kestrel", "Kestrel"), ("Jumping Spider", "JumpingSpider"), ("DragonFly", "Dragonfly"),
    ("HummingBird", "Hummingbird"),
]
_NAMES = [n for n, _ in _ENTRIES]


def animal_classes() -> Dict[str, type]:
    return {n: getattr(A, c) for n, c in _ENTRIES}


def animal_choices() -> List[dict]:
    """Instances are created lazily by the caller's first `visualize` (constructors touch no GPU)."""
    return [{"name": n, "value": cls()} for n, cls in animal_classes().items()]
