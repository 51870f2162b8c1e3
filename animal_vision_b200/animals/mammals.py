"""The 19 mammals that share dog.py's six-step recipe (reference animals/dog.py:14-61 and siblings):
validate -> normalise -> sRGB decode -> 3x3 dichromat collapse -> step-5 filter -> encode.

Step 5 is an isotropic acuity blur (K2), a per-row "streak" blur, or the rat's S-cone row gain
(K1).  Species parameters: SURVEY.md 8a-8, cited per class below.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from .. import tables
from ..engine import get_engine
from .._abi import AVB_F32_GAUSS, AVB_F32_POINT, AVB_F32_STREAK
from .animal import Animal, is_frame, run_single, run_single_float


class _Mammal(Animal):
    ALPHA = 0.6
    S_SCALE = 1.0

    def _matrix(self):
        return tables.dichromat_matrix(self.ALPHA, self.S_SCALE)

    def _run(self, eng, frames, out):
        raise NotImplementedError

    def _run_f32(self, eng, frames, out, tmp, quantize):
        raise NotImplementedError

    def visualize_batch(self, frames, out=None):
        eng = get_engine(frames.device)
        if out is None:
            out = eng.torch.empty_like(frames)
        self._run(eng, frames, out)
        return frames, out

    def visualize(self, image: np.ndarray) -> Optional[Tuple[np.ndarray, np.ndarray]]:
        assert is_frame(image)                       # dog.py:33
        eng = get_engine()
        if image.dtype != np.uint8:                  # float in => float out (dog.py:56-59): the float32 device path
            out = run_single_float(eng, image, lambda d_in, d_out, d_tmp, q: self._run_f32(eng, d_in, d_out, d_tmp, q))
            return image, out
        (out,) = run_single(eng, image, lambda d_in, d_out: self._run(eng, d_in, d_out[0]))
        return image, out                            # baseline is the input object itself (dog.py:61)


class _GaussMammal(_Mammal):
    SIGMA = 1.5

    def _run(self, eng, frames, out):
        taps = tables.gaussian_taps(tables.gaussian_ksize(self.SIGMA), self.SIGMA)
        eng.dichromat_blur(frames, out, self._matrix(), taps)

    def _run_f32(self, eng, frames, out, tmp, quantize):
        taps = tables.gaussian_taps(tables.gaussian_ksize(self.SIGMA), self.SIGMA)
        eng.dichromat_f32(frames, out, tmp, self._matrix(), AVB_F32_GAUSS, taps=taps, quantize=quantize)


class _StreakMammal(_Mammal):
    STREAK = (0.5, 0.8, 2.2, 6.0)    # y_center, sigma_streak, sigma_far, falloff
    CHROMA = 0.0

    def _run(self, eng, frames, out):
        eng.streak_blur(frames, out, self._matrix(), self.STREAK, self.CHROMA)

    def _run_f32(self, eng, frames, out, tmp, quantize):
        eng.dichromat_f32(frames, out, tmp, self._matrix(), AVB_F32_STREAK, streak=self.STREAK, chroma=self.CHROMA, quantize=quantize)


def _gauss(name, alpha, s_scale, sigma, cite):
    return type(name, (_GaussMammal,), {"ALPHA": alpha, "S_SCALE": s_scale, "SIGMA": sigma,
                                        "__doc__": f"{name}: dichromat ({alpha}, {s_scale}) + acuity blur sigma={sigma} ({cite})."})


def _streak(name, alpha, s_scale, streak, cite, chroma=0.0):
    return type(name, (_StreakMammal,), {"ALPHA": alpha, "S_SCALE": s_scale, "STREAK": streak, "CHROMA": chroma,
                                         "__doc__": f"{name}: dichromat ({alpha}, {s_scale}) + streak blur {streak}"
                                                    f"{' + chroma ' + str(chroma) if chroma else ''} ({cite})."})


Dog = _gauss("Dog", 0.58, 0.65, 3.5, "dog.py:46,51")
Bear = _gauss("Bear", 0.6, 0.95, 1.6, "bear.py:29,34")
Lion = _gauss("Lion", 0.6, 0.95, 1.2, "lion.py:29,34")
Tiger = _gauss("Tiger", 0.6, 0.95, 1.2, "tiger.py:29,34")
Elephant = _gauss("Elephant", 0.6, 0.95, 1.8, "elephant.py:29,34")
Fox = _gauss("Fox", 0.65, 0.98, 1.3, "fox.py:29,34")
Wolf = _gauss("Wolf", 0.65, 0.95, 1.4, "wolf.py:29,34")
Raccoon = _gauss("Raccoon", 0.6, 0.98, 2.0, "raccoon.py:29,34")
Squirrel = _gauss("Squirrel", 0.55, 1.05, 0.7, "squirrel.py:29,34")

Cow = _streak("Cow", 0.84, 1.07, (0.5, 0.9, 2.3, 6.5), "cow.py:29,34")
Deer = _streak("Deer", 0.6, 0.95, (0.5, 0.8, 2.6, 8.0), "deer.py:29,34")
Goat = _streak("Goat", 0.75, 1.06, (0.5, 0.8, 2.4, 8.0), "goat.py:29,34")
Horse = _streak("Horse", 0.30, 1.02, (0.5, 0.8, 2.2, 6.0), "horse.py:29,34")
Kangaroo = _streak("Kangaroo", 0.6, 0.98, (0.55, 0.8, 2.3, 8.0), "kangaroo.py:29,34")
Sheep = _streak("Sheep", 0.74, 1.06, (0.48, 0.8, 2.2, 6.0), "sheep.py:30,35")
Panda = _streak("Panda", 0.58, 0.74, (0.52, 1.0, 2.1, 4.5), "panda.py:29-37", chroma=0.06)
Rabbit = _streak("Rabbit", 0.20, 1.01, (0.52, 0.9, 2.5, 5.0), "rabbit.py:29-37", chroma=0.06)
# pig.py:35 drops the blur's return value but the blur mutates its float32 argument in place, so
# the blur applies; pig.py:38 drops apply_chroma_compression's result, so chroma does NOT apply.
Pig = _streak("Pig", 0.89, 1.32, (0.5, 1.2, 2.5, 3.0), "pig.py:30-38")


class Rat(_Mammal):
    """Rat: dichromat (0.05, 0.86) + S-cone vertical gain (rat.py:29,34; animal_utils.py:206-259)."""
    ALPHA, S_SCALE = 0.05, 0.86
    SCONE = dict(s_top=1.3, s_bottom=0.5, power=1.4, extra_boost=0.25)

    def _run(self, eng, frames, out):
        H = frames.shape[1]
        gain = eng.cached(("scone", H, tuple(sorted(self.SCONE.items()))),
                          lambda: eng._dev(tables.scone_row_gain(H, **self.SCONE)))
        eng.colorimetric(frames, out, self._matrix(), row_gain=gain)

    def _run_f32(self, eng, frames, out, tmp, quantize):
        H = frames.shape[1]
        gain = eng.cached(("scone", H, tuple(sorted(self.SCONE.items()))),
                          lambda: eng._dev(tables.scone_row_gain(H, **self.SCONE)))
        eng.dichromat_f32(frames, out, tmp, self._matrix(), AVB_F32_POINT, row_gain=gain, quantize=quantize)


MAMMALS = {c.__name__.lower(): c for c in (Dog, Bear, Lion, Tiger, Elephant, Fox, Wolf, Raccoon, Squirrel, Rat,
                                           Cow, Deer, Goat, Horse, Kangaroo, Sheep, Panda, Rabbit, Pig)}
