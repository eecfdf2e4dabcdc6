"""CPU: the roadmap-builder restatements (oracle/world_oracle.py)."""
import numpy as np

from oracle.world_oracle import (build_neighbors, build_neighbors_bruteforce, philox4x32_10,
                                 sample_states)


def test_philox_known_answer_vectors():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kat:
        got = philox4x32_10(np.array([ctr], np.uint32), key)[0]
        assert tuple(int(v) for v in got) == want


def test_uniform_sampler_ranges_and_determinism():
    a = sample_states(19, 5000, 95.3, "uniform")
    b = sample_states(19, 5000, 95.3, "uniform")
    assert a.dtype == np.float32 and np.array_equal(a, b)
    assert (a[:, :2] >= 0).all() and (a[:, :2] < 95.3).all()
    assert (a[:, 2] >= -np.pi - 1e-6).all() and (a[:, 2] <= np.pi + 1e-6).all()
    assert not np.array_equal(a, sample_states(20, 5000, 95.3, "uniform"))


def test_lattice_sampler_is_the_reference_layout():
    """4 m grid, 7 yaws per location, k = 4 skipped, wrap at 180 (cuda_agent.py:95-109); the interior
    degree is 8 locations x 7 yaws = 56 only because f32 sqrt(32) <= f64 4*sqrt(2)."""
    s = sample_states(0, 7 * 49, 24.0, "lattice")
    assert np.allclose(np.rad2deg(s[:7, 2]), [0, 45, 90, 135, -135, -90, -45], atol=1e-4)
    nbr, num = build_neighbors(s)
    assert num.max() == 56 and num[7 * 24] == 56          # centre location of the 7 x 7 block
    assert num[0] == 21                                   # corner: 3 neighbouring locations


def test_binned_neighbours_equal_bruteforce():
    for seed, n, side in ((1, 700, 40.0), (2, 1500, 52.0), (3, 64, 4.0)):
        s = sample_states(seed, n, side, "uniform")
        a = build_neighbors(s)
        b = build_neighbors_bruteforce(s)
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0])
    s = sample_states(0, 7 * 30, 20.0, "lattice")
    a, b = build_neighbors(s), build_neighbors_bruteforce(s)
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0])
