"""Device key derivation (csrc/prng.cu) vs the oracle's threefry split: bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_key_split_and_chain_bit_exact(built_lib):
    from oracle import prng as oprng
    from to_ued_b200.util import prng
    keys = np.stack([oprng.PRNGKey(s) for s in (0, 1, 42, 2 ** 31 - 1)]).astype(np.uint32)
    kd = prng.to_device(keys, "cuda")
    for num in (1, 2, 3, 7, 512):
        want = np.stack([oprng.split(k, num) for k in keys])
        got = prng.split_device(kd, num).cpu().numpy().view(np.uint32)
        assert np.array_equal(got, want), f"split num={num}"
    # a rank's slice of split(rng, n_global)
    got = prng.split_device(kd[:1], 512, 128, 64).cpu().numpy().view(np.uint32)[0]
    assert np.array_equal(got, oprng.split(keys[0], 512)[128:192])
    # rng, _rng = split(rng) chains (lpg_agent.py:104-105)
    out, carry = prng.chain_device(kd, 5, want_carry=True)
    out, carry = out.cpu().numpy().view(np.uint32), carry.cpu().numpy().view(np.uint32)
    for i, k in enumerate(keys):
        r = k
        for j in range(5):
            r, sub = oprng.split(r, 2)
            assert np.array_equal(out[j, i], sub)
        assert np.array_equal(carry[i], r)
