"""CPU oracle for the animal-vision per-frame pixel pipeline.  TEST INFRASTRUCTURE ONLY.

This package is a NumPy / OpenCV / torch-CPU restatement of the reference's algorithm for the hot
path named in BASELINE.json (`Animal.visualize` for the dichromat mammals, Cat, HoneyBee, and the
MST++ RGB->HSI network).  Every function cites the reference file:line it follows.

Rules (enforced by tests/test_layout.py):
  * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
    import anything from here -- as the checker or as the timed CPU baseline, never as the product.
  * animal_vision_b200/ never imports oracle/ and has no CPU fallback: it raises if the CUDA
    library is missing.

Pinning: the reference ships no tests or golden vectors, so parity is pinned
against the reference ITSELF: tools/make_golden.py imports /root/reference in place (container
only), runs the real `visualize()` / `MST_Plus_Plus.forward`, and commits input seeds + output
arrays / hashes under tests/golden/.  tests/test_oracle_golden.py checks this oracle against
those fixtures (bit-exact for uint8 outputs).  Third-party arithmetic the reference calls and that
is not under /root/reference: OpenCV 4.13.0 (GaussianBlur, remap, resize), NumPy 2.3.5
(power, percentile), torch 2.11.0 (MST++ layers) -- the installed versions are the de-facto oracle
and are recorded inside every fixture.
"""

from . import colorimetry, cvops, mammals, uv  # noqa: F401
