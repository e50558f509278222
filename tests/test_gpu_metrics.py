"""GPU parity of the trajectory-metrics kernels through the C ABI, via the reference-shaped
MetricsCalculator / IntersectionVolumeGuide classes: against fixtures produced by the unmodified reference
(tests/golden/metrics.npz, oracle/make_golden_metrics.py) and against the CPU oracle on seeded ensembles."""
import numpy as np
import pytest
import torch

from oracle import metrics_oracle as mo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
# float64 arithmetic on the device; the reference's end-effector chain is float32 (FK, norms, and numpy's float32 FFT),
# so end-effector quantities agree to float32 round-off: 2e-6 m on positions, 1e-5 relative on SPARC
POS_TOL, SAL_RTOL = 2e-6, 1e-5


@pytest.fixture(scope="module")
def calc():
    from edmp_b200 import IntersectionVolumeGuide
    from edmp_b200.lib import MetricsCalculator
    guide = IntersectionVolumeGuide(obstacle_config=np.array([[0.5, 0, 0.3, 0, 0, 0, 1, 0.1, 0.1, 0.1]]), device=DEV,
                                    guide_cfgs={}, batch_size=1)
    return MetricsCalculator(guide)


def test_ee_transform_matches_reference_fixture(golden, calc):
    g = golden("metrics.npz")
    q = calc.guide.rearrange_joints(torch.tensor(g["traj"], dtype=torch.float32, device=DEV))
    T = calc.guide.get_end_effector_transform(q)
    assert T.shape == (g["traj"].shape[0], 50, 4, 4) and T.dtype == torch.float32
    np.testing.assert_allclose(T.cpu().numpy(), g["ee_transforms"], rtol=0, atol=POS_TOL)


def test_metrics_match_reference_fixture(golden, calc):
    g = golden("metrics.npz")
    traj, dts = g["traj"], g["dts"]
    for i, dt in enumerate(dts):
        r = calc.ensemble_metrics(traj, float(dt))
        np.testing.assert_allclose(r["joint_path_length"], g["path_lengths"][:, 0], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(r["end_eff_path_length"], g["path_lengths"][:, 1], rtol=0, atol=50 * POS_TOL)
        np.testing.assert_allclose(r["joint_smoothness"], g["sparc"][i, :, 0], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(r["end_eff_smoothness"], g["sparc"][i, :, 1], rtol=SAL_RTOL, atol=1e-7)
    assert r["joint_smoothness"][-1] == 0.0 and r["end_eff_smoothness"][-1] == 0.0   # constant trajectory (:86-88)


def test_reference_api_shapes(golden, calc):
    g = golden("metrics.npz")
    js, es = calc.smoothness_metric(g["traj"][0], float(g["dts"][0]))
    sal, (f, Mf), (f_sel, Mf_sel) = js
    np.testing.assert_allclose(sal, g["sparc"][0, 0, 0], rtol=1e-9)
    np.testing.assert_array_equal(f, g["f"])
    np.testing.assert_allclose(Mf, g["Mf_joint"], rtol=0, atol=1e-12)
    np.testing.assert_array_equal(f_sel, g["fsel_joint"])
    assert Mf_sel.shape == f_sel.shape
    np.testing.assert_allclose(es[1][1], g["Mf_ee"], rtol=0, atol=2e-6)
    jl, el = calc.path_length_metric(g["traj"][0])
    np.testing.assert_allclose([jl, el], g["path_lengths"][0], rtol=1e-5)
    # the known-answer example of the reference's docstring (lib/metrics.py:79-84)
    ex = calc.sparc(g["example_move"], fs=100.)
    assert "%.5f" % ex[0] == "-1.41403"
    np.testing.assert_allclose(ex[0], float(g["example_sal"]), rtol=1e-10)
    assert calc.sparc(np.zeros(49), 50.) == (0, None, None)
    assert calc.smoothness_metric(g["traj"][-1], 0.02) == ((0, None, None), (0, None, None))


def test_metrics_match_oracle_on_a_seeded_ensemble(calc):
    """1024 random-walk trajectories (ragged lengths too): every row against the CPU oracle on a sample, and the
    batch against itself row by row (rows are independent)."""
    rng = np.random.default_rng(3)
    for n in (50, 17, 64):
        traj = np.cumsum(rng.normal(scale=0.03, size=(1024, 7, n)), axis=2) + rng.uniform(-1, 1, size=(1024, 7, 1))
        r = calc.ensemble_metrics(traj, 0.04)
        for row in (0, 1, 511, 1023):
            m = mo.trajectory_metrics(traj[row], 0.04)
            got = [r[k][row] for k in ("joint_path_length", "end_eff_path_length", "joint_smoothness", "end_eff_smoothness")]
            np.testing.assert_allclose(got, m, rtol=SAL_RTOL, atol=1e-6)
        one = calc.ensemble_metrics(traj[700:701], 0.04)
        for k in one:
            assert one[k][0] == r[k][700]


def test_metrics_reject_bad_arguments(calc):
    from edmp_b200._lib import EdmpError
    with pytest.raises(ValueError):
        calc.ensemble_metrics(np.zeros((2, 6, 50)), 0.1)
    with pytest.raises(EdmpError):
        calc.ensemble_metrics(np.zeros((2, 7, 50)), 0.0)
    with pytest.raises(EdmpError):
        calc.ensemble_metrics(np.zeros((2, 7, 65)), 0.1)
    with pytest.raises(EdmpError):
        calc.ensemble_metrics(np.zeros((2, 7, 50)), 0.1, padlevel=9)   # nfft = 32768 > 8192
