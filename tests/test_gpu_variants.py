"""GPU parity of the measured variants behind the run-time switches (INTEGRATION.md "Run-time switches").

The switches are read once per process, so every variant runs in a subprocess.  Two kinds of claims:
* HoneyBee's two map routes (planes of the hist pass / second walk) give IDENTICAL bytes for every mapper with percentiles;
* every MST++ schedule (fused / three-kernel feed-forward block, attention side kernel / three launches, fused kernel also at
  level 1) stays inside the 1e-2 gate against the fp32 oracle, and the schedules agree with each other far inside it."""
import hashlib
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

UV_CHILD = r"""
import hashlib, json, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import torch
import animal_vision_b200.animals as A
g = torch.Generator().manual_seed(5)
out = []
for shape in ((3, 270, 480), (1, 123, 236), (2, 64, 1100), (1, 37, 50)):
    fr = torch.randint(0, 256, (*shape, 3), dtype=torch.uint8, generator=g).cuda()
    fr[0, : shape[1] // 2] //= 3
    for kw in ({}, {"adaptation": "gray_world"}, {"blur_sigma_px": None}, {"blur_sigma_px": 0.6}, {"adaptation": None},
               {"mapping_mode": "falsecolor"}, {"mapping_mode": "uv_purple_yellow"}, {"mapping_mode": "falsecolor_uv_mixed"},
               {"mapping_mode": "falsecolor", "adaptation": "gray_world", "blur_sigma_px": 0.6}):
        res = A.HoneyBee(**kw).visualize_batch(fr)
        res = res[-1] if isinstance(res, (tuple, list)) else res
        out.append(hashlib.sha1(res.cpu().numpy().tobytes()).hexdigest())
    # a strided view (rows not packed) must give the packed result
    wide = torch.zeros((shape[0], shape[1], shape[2] + 24, 3), dtype=torch.uint8, device="cuda")
    wide[:, :, 8:8 + shape[2]] = fr
    res = A.HoneyBee().visualize_batch(wide[:, :, 8:8 + shape[2]])
    res = res[-1] if isinstance(res, (tuple, list)) else res
    out.append(hashlib.sha1(res.contiguous().cpu().numpy().tobytes()).hexdigest())
print("RESULT " + json.dumps(out))
"""

K4_CHILD = r"""
import json, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch
from oracle import mstpp as O
from animal_vision_b200.mstpp import MSTPlusPlus
sd = O.make_weights(0)
net = MSTPlusPlus(sd)
res = {}
for (b, h, w) in ((1, 64, 72), (2, 24, 200), (1, 130, 260)):
    x = torch.rand(b, 3, h, w, generator=torch.Generator().manual_seed(10 + h))
    ref = O.forward(x, sd).numpy()
    y = net(x.cuda()).cpu().numpy()
    res["%%dx%%dx%%d" %% (b, h, w)] = {"max": float(np.abs(y - ref).max() / np.abs(ref).max()),
                                  "l2": float(np.linalg.norm(y - ref) / np.linalg.norm(ref)),
                                  "sum": float(np.float64(y).sum()), "abs": float(np.abs(ref).max())}
np.save(%(out)r, y)
print("RESULT " + json.dumps(res))
"""


def _child(code, env):
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


def test_honeybee_plane_route_equals_the_second_walk():
    a = _child(UV_CHILD % {"root": ROOT}, {"AVB_UV_NO_PLANE_MAP": "0"})
    b = _child(UV_CHILD % {"root": ROOT}, {"AVB_UV_NO_PLANE_MAP": "1"})
    assert len(a) == len(b) == 40
    assert a == b


@pytest.mark.parametrize("env", [
    {"AVB_MSTPP_FFN_UNFUSED": "1"},
    {"AVB_MSTPP_ATTN_UNMERGED": "1"},
    {"AVB_MSTPP_ATTN_UNMERGED": "1", "AVB_MSTPP_STATS_CUDA_CORES": "1", "AVB_MSTPP_FFN_UNFUSED": "1"},
    {"AVB_MSTPP_FFN_FUSED_MAXCP": "64"},
    {"AVB_MSTPP_DW_WALK": "1"},
])
def test_mstpp_schedules_agree(env, tmp_path):
    import numpy as np
    base_npy, var_npy = str(tmp_path / "base.npy"), str(tmp_path / "var.npy")
    base = _child(K4_CHILD % {"root": ROOT, "out": base_npy}, {})
    var = _child(K4_CHILD % {"root": ROOT, "out": var_npy}, env)
    for key in base:
        for r in (base[key], var[key]):
            assert r["max"] <= 1e-2 and r["l2"] <= 1e-2, (env, key, r)
    yb, yv = np.load(base_npy), np.load(var_npy)
    scale = np.abs(yb).max()
    assert np.abs(yb - yv).max() / scale <= 5e-3, (env, float(np.abs(yb - yv).max() / scale))
