"""CPU checks on the cached full-size oracle sweeps (tests/golden/fullsize_*.npz): the two oracle arithmetics (fp32 = the
reference's, bf16 GEMM inputs = the engine's contract) agree within the north_star tolerance on every window, and the top-k
groups derived from them agree as SETS (the order inside a group is only defined up to the arithmetic noise)."""
from pathlib import Path

import numpy as np
import pytest

from audio_deepfake_explainability_b200 import grid

GOLDEN = Path(__file__).parent / "golden"


@pytest.mark.parametrize("family", ["REAL", "SUNO"])
def test_fp32_and_bf16_oracles_agree_on_every_window(family):
    z = np.load(GOLDEN / f"fullsize_occlusion_{family}0.npz")
    a, b = z["delta_fp32"], z["delta_bf16"]
    assert a.shape == b.shape == (228,) and z["windows"].shape == (228, 4)
    err = float(np.abs(a - b).max())
    assert err < 1e-3 and abs(float(z["base_fp32"]) - float(z["base_bf16"])) < 1e-3
    margins = grid.topk_boundary_margins(a, 5)
    ga, gb = grid.topk_window_groups(a, 5), grid.topk_window_groups(b, 5)
    for g in ("best", "worst", "most_influential"):
        if margins[g] > 2 * err:
            assert set(ga[g].tolist()) == set(gb[g].tolist())


def test_margins_and_snapping():
    d = np.array([0.5, -0.4, 3e-6, 0.0, -2e-6, 0.3, 1e-3])
    m = grid.topk_boundary_margins(d, 2)
    assert np.isclose(m["best"], 0.4 - 0.3) and np.isclose(m["worst"], 3e-6 - 2e-6)
    assert grid.topk_window_groups(d, 2)["worst"].tolist() == [3, 4]
    # snapped: 2, 3, 4 are exact ties -> grid order
    assert grid.topk_window_groups(d, 2, 1e-4)["worst"].tolist() == [2, 3]
    assert grid.topk_window_groups(d, 2, 1e-4)["best"].tolist() == [0, 1]
    assert grid.topk_boundary_margins(d, 10)["best"] == float("inf")
