"""Host side of K4 (CPU): the parameter order handed to avb_mstpp_create is the reference's
state_dict order, and the oracle's functional network is pinned to the reference's nn.Module."""
import numpy as np
import torch

from oracle import mstpp as O


def test_param_order_matches_oracle_and_counts():
    from animal_vision_b200 import mstpp
    a, b = mstpp.param_order(), O.param_shapes()
    assert list(a.items()) == list(b.items())
    assert len(a) == 227
    assert sum(int(np.prod(s)) for s in a.values()) == mstpp.N_PARAMS == 1619625


def test_flatten_state_dict_layout_and_module_prefix():
    from animal_vision_b200 import mstpp
    sd = O.make_weights(0)
    flat = mstpp.flatten_state_dict(sd)
    assert flat.dtype == np.float32 and flat.size == mstpp.N_PARAMS
    assert np.array_equal(flat[:31 * 27], sd["conv_in.weight"].numpy().ravel())
    assert np.array_equal(flat[-31 * 31 * 9:], sd["conv_out.weight"].numpy().ravel())
    prefixed = {"module." + k: v for k, v in sd.items()}
    assert np.array_equal(mstpp.flatten_state_dict(prefixed), flat)


def test_oracle_forward_matches_reference_golden(golden, golden_meta):
    x = torch.rand(1, 3, 42, 52, generator=torch.Generator().manual_seed(1))
    y = O.forward(x, O.make_weights(0)).numpy()
    ref = golden("mstpp")["y"]
    assert np.abs(y - ref).max() <= 2e-5 * np.abs(ref).max()


def test_synthetic_state_dict_equals_oracle_weights():
    """bench.py / tools take their seeded weights from the product (nothing product-side imports oracle/)."""
    from animal_vision_b200 import mstpp
    a, b = mstpp.synthetic_state_dict(0), O.make_weights(0)
    assert list(a) == list(b)
    assert all(torch.equal(a[k], b[k]) for k in a)
