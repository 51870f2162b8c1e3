"""Oracle: sRGB transfer functions, frame normalisation and the LMS cone matrices.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates reference animals/animal_utils.py.
dtype behaviour is part of the contract: NumPy's type promotion decides where the reference
silently switches from float32 to float64, and the oracle keeps those switches.
"""
from __future__ import annotations

import numpy as np

# animals/animal_utils.py:56-62 (float32) and :70-75 (float64: no dtype given)
RGB_TO_LMS = np.array(
    [[0.31399022, 0.63951294, 0.04649755],
     [0.15537241, 0.75789446, 0.08670142],
     [0.01775239, 0.10944209, 0.87256922]], dtype=np.float32)
LMS_TO_RGB = np.array(
    [[5.472213, -4.6419606, 0.16963711],
     [-1.125242, 2.2931712, -0.16789523],
     [0.02980164, -0.19318072, 1.1636479]])  # float64 on purpose


def decode_srgb(v: np.ndarray) -> np.ndarray:
    """IEC 61966-2-1 EOTF, both branches evaluated (animal_utils.py:5-11). dtype follows input."""
    return np.where(v <= 0.04045, v / 12.92, ((v + 0.055) / (1 + 0.055)) ** 2.4)


def encode_srgb(v: np.ndarray) -> np.ndarray:
    """IEC 61966-2-1 OETF (animal_utils.py:13-19). Callers clip to [0,1] first."""
    return np.where(v <= 0.0031308, 12.92 * v, (1 + 0.055) * (v ** (1 / 2.4)) - 0.055)


def is_frame(a) -> bool:
    """animal_utils.py:21-39: ndarray, HxWx3, numeric dtype."""
    return (isinstance(a, np.ndarray) and a.ndim == 3 and a.shape[2] == 3
            and np.issubdtype(a.dtype, np.number))


def normalize_frame(a: np.ndarray) -> np.ndarray:
    """animal_utils.py:41-50: float32 copy; divide by 255 ONLY IF the frame max exceeds 1.0."""
    f = a.astype(np.float32)
    if f.max() > 1.0:
        f /= 255.0
    return np.clip(f, 0.0, 1.0)


def divides_by_255(a: np.ndarray) -> bool:
    """The data-dependent branch of normalize_frame, exposed for the tests."""
    return bool(a.astype(np.float32).max() > 1.0)


def dichromat_matrix(alpha: float, s_scale: float) -> np.ndarray:
    """animal_utils.py:88-119. T such that the reference applies `pixels @ T.T`.

    basis(f32 eye) @ RGB_TO_LMS.T -> f32 ; @ D.T (f32) -> f32 ; @ LMS_TO_RGB.T (f64) -> f64 ; cast f32.
    """
    basis_lms = np.eye(3, dtype=np.float32) @ RGB_TO_LMS.T
    D = np.array([[alpha, 1.0 - alpha, 0.0],
                  [alpha, 1.0 - alpha, 0.0],
                  [0.0, 0.0, s_scale]], dtype=np.float32)
    return ((basis_lms @ D.T) @ LMS_TO_RGB.T).astype(np.float32)


def apply_matrix(lin: np.ndarray, T: np.ndarray) -> np.ndarray:
    """dog.py:44-48: flatten to (N,3), right-multiply by T.T, restore shape."""
    return (lin.reshape(-1, 3) @ T.T).reshape(lin.shape)


def quantize(srgb01: np.ndarray, dtype) -> np.ndarray:
    """dog.py:56-59: integer dtypes get x*255+0.5 truncated; float dtypes are a plain cast."""
    if np.issubdtype(dtype, np.integer):
        return (srgb01 * 255.0 + 0.5).astype(dtype)
    return srgb01.astype(dtype)


def encode_tail(lin: np.ndarray, dtype) -> np.ndarray:
    """dog.py:54-59: clip -> OETF -> clip -> quantise."""
    return quantize(np.clip(encode_srgb(np.clip(lin, 0.0, 1.0)), 0.0, 1.0), dtype)


# ---- host tables the CUDA path also uses; rebuilt here independently for the tests ----

def decode_lut_u8(div255: bool = True) -> np.ndarray:
    """256-entry float32 table: decode_srgb(normalize(v)) for v = 0..255.

    With div255=False (frame max <= 1) values are NOT divided: normalize_frame clips them to [0,1],
    so entry v is decode(min(v,1)).
    """
    v = np.arange(256, dtype=np.float32)
    x = np.clip(v / np.float32(255.0) if div255 else v, 0.0, 1.0).astype(np.float32)
    return decode_srgb(x).astype(np.float32)
