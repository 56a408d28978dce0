"""jax.random restated: Threefry-2x32 known answers (Random123 KAT vectors, the values JAX's own tests use) and the
jax.random.split(PRNGKey(0)) key data of both key layouts."""
import numpy as np

from open_duck_playground_b200 import rng as jr


def test_threefry_known_answers():
    kat = [((0, 0), (0, 0), (0x6B200159, 0x99BA4EFE)),
           ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
           ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))]
    for key, ctr, exp in kat:
        out = jr.threefry2x32(key[0], key[1], ctr[0], ctr[1])
        assert (int(out[0]), int(out[1])) == exp


def test_split_key0_original_layout():
    # jax.random.split(PRNGKey(0)) with jax_threefry_partitionable=False: counts [0,1,2,3] -> x0=[0,1], x1=[2,3]
    y0, y1 = jr.threefry2x32(0, 0, np.array([0, 1], np.uint32), np.array([2, 3], np.uint32))
    assert y0.tolist() == [4146024105, 967050713] and y1.tolist() == [2718843009, 1272950319]


def test_split_key0_partitionable_layout():
    # the JAX >= 0.5 default the reference pins (pyproject.toml:8)
    assert jr.split(jr.PRNGKey(0), 2).tolist() == [[1797259609, 2579123966], [928981903, 3453687069]]


def test_uniform_range_and_determinism():
    k = jr.split(jr.PRNGKey(7), 3)[1]
    u = jr.uniform(k, 1000, -0.05, 0.05)
    assert u.dtype == np.float32 and u.min() >= -0.05 and u.max() < 0.05
    assert np.array_equal(u, jr.uniform(k, 1000, -0.05, 0.05))
    assert abs(float(u.mean())) < 5e-3
