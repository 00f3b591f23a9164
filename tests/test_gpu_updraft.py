"""Stage 1 on the GPU vs the oracle and the reference's golden outputs (through the C-ABI)."""
import os

import numpy as np
import pytest

from oracle import oracle_np as O

pytestmark = pytest.mark.gpu

TOL = 1e-5      # north_star: fields within 1e-5 relative (float32 device arithmetic vs float64 reference)


def rel(a, b):
    return np.abs(np.asarray(a, dtype=np.float64) - b).max() / max(np.abs(b).max(), 1e-30)


def check_fields(out, sl, asp, oro, K):
    assert rel(out["slope"], sl) <= TOL
    # aspect is compared away from the dz_dx == 0 convention cells (they agree too when fp32 sums are exact)
    assert rel(out["aspect"], asp) <= TOL
    assert rel(out["orograph"], oro) <= TOL
    assert rel(out["updraft"], K) <= TOL
    for k in out:
        a = out[k]
        assert (a[0] == 0).all() and (a[-1] == 0).all() and (a[:, 0] == 0).all() and (a[:, -1] == 0).all()


@pytest.mark.parametrize("name", ["a", "b"])
def test_golden(golden, name):
    from ssrs_b200 import layers
    g = golden("stencil")
    z, res = g[f"{name}_z"], float(g[f"{name}_res"])
    out = layers.updraft_fields(z, res, 10.0, 270.0, 0.75)
    check_fields(out, g[f"{name}_slope"], g[f"{name}_aspect"], g[f"{name}_oro"], g[f"{name}_K"])
    out = layers.updraft_fields(z, res, g[f"{name}_ws"], g[f"{name}_wd"], 0.75, want=("orograph", "updraft"))
    assert rel(out["orograph"], g[f"{name}_oro_cell"]) <= TOL
    assert rel(out["updraft"], g[f"{name}_K_cell"]) <= TOL
    # the reference-named single-output functions
    assert rel(layers.compute_slope_degrees(z, res), g[f"{name}_slope"]) <= TOL
    assert rel(layers.compute_aspect_degrees(z, res), g[f"{name}_aspect"]) <= TOL
    oro = layers.compute_orographic_updraft(10.0, 270.0, g[f"{name}_slope"], g[f"{name}_aspect"])
    assert rel(oro, g[f"{name}_oro"]) <= TOL
    assert rel(layers.get_above_threshold_speed(g[f"{name}_oro"], 0.75), g[f"{name}_K"]) <= TOL


@pytest.mark.parametrize("shape,res", [((500, 600), 100.0), ((3, 3), 10.0), ((5, 1000), 10.0), ((257, 131), 10.0),
                                       ((1024, 1280), 10.0)])
@pytest.mark.parametrize("path", ["auto", "plain", "tma"])
def test_oracle_sizes(shape, res, path, monkeypatch):
    from ssrs_b200 import layers
    from ssrs_b200.synth import synthetic_dem
    rows, cols = shape
    if path == "tma" and cols % 4:
        pytest.skip("TMA staging needs a 16-byte pitch")
    if path != "auto":
        monkeypatch.setenv("SSRS_STENCIL_PATH", path)
    z = synthetic_dem(rows, cols, res, seed=rows + cols)
    sl, asp, oro, K = O.updraft_pipeline(z, res, 10.0, 270.0, 0.75)
    out = layers.updraft_fields(z, res, 10.0, 270.0, 0.75)
    check_fields(out, sl, asp, oro, K)
    # oblique wind, different threshold
    sl, asp, oro, K = O.updraft_pipeline(z, res, 7.5, 33.0, 0.5)
    out = layers.updraft_fields(z, res, 7.5, 33.0, 0.5)
    check_fields(out, sl, asp, oro, K)


def test_full_size_properties():
    """BASELINE config 2 grid (5000, 6000) at 10 m: checked through size-independent properties —
    a random sample of windows against the oracle, linearity in wind speed, zero border, staging-path
    agreement (TMA vs plain must be bit-identical: same arithmetic, different loader)."""
    import torch
    from ssrs_b200 import layers
    from ssrs_b200.synth import synthetic_dem
    rows, cols, res = 5000, 6000, 10.0
    z = synthetic_dem(rows, cols, res)
    zt = torch.from_numpy(z).cuda()
    out = layers.updraft_fields(zt, res, 10.0, 270.0, 0.75)
    rng = np.random.RandomState(0)
    for _ in range(6):
        r0, c0 = rng.randint(0, rows - 300), rng.randint(0, cols - 300)
        win = z[r0:r0 + 300, c0:c0 + 300]
        sl, asp, oro, K = O.updraft_pipeline(win, res, 10.0, 270.0, 0.75)
        sub = {k: v[r0 + 1:r0 + 299, c0 + 1:c0 + 299].cpu().numpy() for k, v in out.items()}
        assert rel(sub["slope"], sl[1:-1, 1:-1]) <= TOL and rel(sub["aspect"], asp[1:-1, 1:-1]) <= TOL
        assert rel(sub["orograph"], oro[1:-1, 1:-1]) <= TOL and rel(sub["updraft"], K[1:-1, 1:-1]) <= TOL
    out2 = layers.updraft_fields(zt, res, 20.0, 270.0, 0.75, want=("orograph",))
    assert torch.allclose(out2["orograph"], 2.0 * out["orograph"], rtol=1e-6, atol=0)
    os.environ["SSRS_STENCIL_PATH"] = "plain"
    try:
        out3 = layers.updraft_fields(zt, res, 10.0, 270.0, 0.75)
    finally:
        del os.environ["SSRS_STENCIL_PATH"]
    for k in out:
        assert torch.equal(out[k], out3[k])
    frac0 = float((out["updraft"] == 0).float().mean())
    assert 0.3 < frac0 < 0.7            # the high-contrast regime the solver must handle


def test_bad_arguments():
    from ssrs_b200 import layers
    with pytest.raises(ValueError):
        layers.updraft_fields(np.zeros((2, 5), np.float32), 10.0, 1.0, 0.0)
    with pytest.raises(ValueError):
        layers.updraft_fields(np.zeros((5, 5), np.float32), -1.0, 1.0, 0.0)
