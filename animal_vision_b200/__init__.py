"""animal_vision_b200 -- B200-native (sm_100a) implementation of animal-vision's per-frame pixel
pipeline behind the reference's Animal / Renderer plugin API.  See DESIGN.md."""
__version__ = "0.1.0"
