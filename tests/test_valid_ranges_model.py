"""Host-side model of the persistent kernel's valid-sample range pre-pass (dsp_slam_b200/csrc/dspgn_solve.cuh:
valid_sample_ranges).  loss.py:68 keeps the ray samples inside the unit sphere; the device enumerates, per ray, only the hull
[first valid, last valid] of its D samples and finds it by testing the samples next to the closed-form chord of the ray in
the unit ball instead of all D.  The model restates both searches with the same per-sample test and checks, on random and
grazing rays, depth ranges that cut the chord, and arbitrary rotations, that the windowed search returns the hull of the
exhaustive one (the GPU tests check the device code itself bit for bit: test_valid_sample_hulls_equal_full_ray_enumeration)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import valid_ranges_model as M


def test_windowed_hull_search_equals_exhaustive_search():
    bad, nonempty, stats = M.run(1500, seed=7)
    assert bad == 0
    assert 400 < nonempty < 1300                      # the sample has hits, misses and grazing rays
    assert stats["tests"] < 8 * 1500                  # a handful of per-sample tests per ray instead of D = 50
