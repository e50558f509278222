"""CPU: the oracle restatement (oracle/) pinned against fixtures produced by the unmodified
reference (tests/golden/, made by oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import guide_oracle as go, guide_params, sampler_oracle as so, scenes, unet_oracle, weights


@pytest.fixture(scope="module")
def sd():
    return weights.seeded_state_dict(0)


def test_state_dict_layout():
    table = weights.key_table()
    assert len(table) == 290
    assert sum(int(np.prod(s)) for _, s, _ in table) == 29938471


def test_unet_oracle_matches_reference(golden, sd):
    g = golden("unet_forward.npz")
    x = torch.tensor(g["x"])
    for t in (255, 128, 1):
        taps = {}
        with torch.no_grad():
            eps = unet_oracle.unet_forward(sd, x, t, taps).numpy()
        np.testing.assert_allclose(eps, g["eps_t%d" % t], rtol=0, atol=1e-6)
        if t == 128:
            for key in g.files:
                if key.startswith("tap_t128/"):
                    ref = g[key]
                    mine = taps[key.split("/", 1)[1]].numpy()
                    np.testing.assert_allclose(mine, ref[:, :, :mine.shape[2]], rtol=0, atol=2e-6)


@pytest.mark.parametrize("case", ["mixed", "iv", "sv_axis"])
def test_guide_oracle_matches_reference(golden, case):
    g = golden("guide.npz")
    guides, bpg = [int(v) for v in g[case + "/guides"]], int(g[case + "/bpg"])
    cfgs = so.expand_guide_tables([guide_params.GUIDES[n] for n in guides], bpg)
    scene, q, ld = g[case + "/scene"], g[case + "/q"], g["link_dims"]
    np.testing.assert_allclose(ld, go.LINK_DIMS, rtol=0, atol=1e-7)
    B = q.shape[0]
    for t in (254, 100, 6):
        ref = g["%s/grad_t%d" % (case, t)]
        scale = max(1.0, np.abs(ref).max())
        a = go.gradient_autograd(q, scenes.START, scenes.GOAL, scene, cfgs, t)
        b = go.gradient_analytic(q, scenes.START, scenes.GOAL, scene, cfgs, t)
        assert np.abs(a - ref).max() <= 2e-6 * scale
        assert np.abs(b - ref).max() <= 2e-6 * scale
    omin, omax = go.obstacle_aabbs(scene, cfgs["expansion"][:, 99], cfgs["clearance"][:, 99], rows=B)
    qt = torch.tensor(q, dtype=torch.float32)
    np.testing.assert_allclose(go.iv_cost(qt, omin, omax).numpy(), g[case + "/iv_t100"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(go.sv_cost(qt, scenes.START, scenes.GOAL, omin, omax).numpy(),
                               g[case + "/sv_t100"], rtol=0, atol=1e-7)
    omin0, omax0 = go.obstacle_aabbs(scene, rows=B)
    np.testing.assert_allclose(go.iv_cost(qt[:, :, :1], omin0, omax0).numpy(), g[case + "/cost_t0"],
                               rtol=0, atol=1e-7)
    traj = np.concatenate([np.broadcast_to(scenes.START[None, :, None], (B, 7, 1)), q,
                           np.broadcast_to(scenes.GOAL[None, :, None], (B, 7, 1))], axis=2)
    fs = go.final_sv_costs(traj, scenes.START, scenes.GOAL, scene)
    np.testing.assert_allclose(fs, g[case + "/final_sv"], rtol=1e-5, atol=1e-7)
    assert int(np.argmin(fs)) == int(g[case + "/best_index"])


def test_guide_nan_poisoning(golden):
    g = golden("guide.npz")
    cfgs = so.expand_guide_tables([guide_params.GUIDES[n] for n in (1, 9)], 1)
    for fn in (go.gradient_autograd, go.gradient_analytic):
        out = fn(g["nan/q"], scenes.START, scenes.GOAL, g["nan/scene"], cfgs, 100)
        assert np.isnan(out).all() and np.isnan(g["nan/grad_t100"]).all()


def test_schedule_and_tables():
    beta, alpha, abar = so.schedule()
    assert beta.shape == (255,) and abs(beta[0] - 0.02 / 255) < 1e-15 and abs(beta[-1] - 0.02) < 1e-15
    assert abs(abar[-1] - np.prod(1 - beta)) < 1e-15
    cfgs = so.expand_guide_tables([guide_params.GUIDES[n] for n in (18, 1)], 2)
    # guide18: isr2 [10,40) then isr3 [0,20) overwrites 10..19 with zeros
    assert cfgs["expansion"][0, 15] == 0.0 and cfgs["expansion"][0, 25] > 0.0
    assert cfgs["expansion"][0, 100] == 0.4 and cfgs["expansion"][2, 100] == 0.0
    assert cfgs["guidance_schedule"][0, 7] == 0.05 and abs(cfgs["guidance_schedule"][2, 254] - (1.4 + 254 / 255)) < 1e-12


def _replay(name):
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "sampler_%s.npz" % name))
    guides, bpg = [int(v) for v in g["guides"]], int(g["bpg"])
    cfgs = so.expand_guide_tables([guide_params.GUIDES[n] for n in guides], bpg)
    B = cfgs["total_batch_size"]
    rng = np.random.default_rng(int(g["noise_seed"]))
    noise = [rng.normal(size=(B, 7, 50)) for _ in range(255)]
    assert abs(sum(n.sum() for n in noise) - float(g["noise_checksum"])) < 1e-9
    return g, cfgs, noise


@pytest.fixture(scope="module")
def sd02():
    return weights.seeded_state_dict(0, final_gain=0.2)


def test_sampler_oracle_teacher_forced(golden, sd02):
    """Single steps from the reference's own recorded states: posterior, guide, update."""
    sd = sd02
    g, cfgs, noise = _replay("mixed")
    beta, alpha, abar = so.schedule()
    for t in [int(s) for s in g["steps"]]:
        x = g["x_in_t%d" % t]
        with torch.no_grad():
            eps = unet_oracle.unet_forward(sd, torch.tensor(x, dtype=torch.float32), t).numpy()
        np.testing.assert_allclose(eps, g["eps_t%d" % t], rtol=0, atol=2e-6)
        xp = so.posterior_step(x, t, g["eps_t%d" % t], noise[255 - t], beta, alpha, abar)
        np.testing.assert_allclose(xp, g["x_post_t%d" % t], rtol=0, atol=1e-12)
        if t % 2 == 0 and t >= 5:
            G = go.gradient_analytic(so.clip_joints(xp[:, :, 1:-1]), scenes.START, scenes.GOAL,
                                     g["scene"], cfgs, t)
            ref = g["grad_t%d" % t]
            assert np.abs(G - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())
            xp[:, :, 1:-1] -= cfgs["guidance_schedule"][:, t - 1, None, None] * ref
        xp[:, :, 0], xp[:, :, -1] = scenes.START, scenes.GOAL
        np.testing.assert_allclose(xp, g["x_out_t%d" % t], rtol=0, atol=1e-12)


def test_sampler_oracle_end_to_end_iv(golden, sd02):
    """Full 255 steps, iv guides [1,2,3] (BASELINE config 3 ensemble): <= 1e-4 rad."""
    sd = sd02
    g, cfgs, noise = _replay("iv")
    out = so.denoise_guided(sd, g["scene"], cfgs, scenes.START, scenes.GOAL, g["x_T"], noise,
                            gradient="analytic")
    assert np.abs(out - g["final"]).max() <= 1e-4
    best = go.choose_best_trajectory(out, scenes.START, scenes.GOAL, g["scene"])
    assert np.abs(best - g["best"]).max() <= 1e-4


# ---- sphere / signed-distance guide family (SURVEY.md section 8 a-S) -----------------------------------------------
@pytest.mark.parametrize("case", ["general", "yaw"])
def test_sdf_oracle_matches_reference_geometry(golden, case):
    """box / cylinder SDF restatement against the reference's mpinets.geometry (fixture made by make_golden_sdf)."""
    from oracle import sdf_oracle as sdfo
    g = golden("sdf.npz")
    pts = torch.tensor(g[case + "/points"])
    box = sdfo.box_sdf(pts, g[case + "/boxes"]).min(dim=-1).values.numpy()
    cyl = sdfo.cylinder_sdf(pts, g[case + "/cylinders"]).min(dim=-1).values.numpy()
    np.testing.assert_allclose(box, g[case + "/cuboid_sdf"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(cyl, g[case + "/cylinder_sdf"], rtol=0, atol=1e-12)


def test_sdf_oracle_fk_matches_the_dh_chain():
    """URDF-chain FK of the sphere guide against the reference-pinned DH chain of lib/guide.py: the seven joint frames
    (origin and all three axes) coincide, so both guide families see the same robot."""
    from oracle import sdf_oracle as sdfo
    q = np.random.default_rng(3).uniform(-2.5, 2.5, size=(16, 7))
    dh = go._fk_frames64(q)
    fk = sdfo.franka_fk(torch.tensor(q))
    for i in range(7):
        np.testing.assert_allclose(fk["panda_link%d" % (i + 1)].numpy(), dh[:, i], rtol=0, atol=1e-12)
    c, r = sdfo.sphere_centres(torch.tensor(q))
    assert c.shape == (16, 59, 3) and r.shape == (59,) and len(set(r.tolist())) == 11


def test_sdf_oracle_gradient_is_consistent():
    """autograd gradient of the oracle cost against central differences"""
    from oracle import sdf_oracle as sdfo, make_golden_sdf
    rng = np.random.default_rng(5)
    boxes, cyls = make_golden_sdf.random_scene(rng, 6, 2)
    q = scenes.START[None, :, None] + (scenes.GOAL - scenes.START)[None, :, None] * np.linspace(0, 1, 6)[None, None, :] \
        + 0.1 * rng.normal(size=(2, 7, 6))
    cost, grad, clr = sdfo.evaluate(q, boxes, cyls, margin=0.05)
    assert cost.max() > 0 and clr.shape == (2, 6)
    h = 1e-6
    for (b, j, w) in [(0, 0, 1), (1, 3, 4), (0, 6, 2), (1, 1, 0)]:
        qp, qm = q.copy(), q.copy()
        qp[b, j, w] += h
        qm[b, j, w] -= h
        fd = (sdfo.evaluate(qp, boxes, cyls, 0.05, False)[0][b] - sdfo.evaluate(qm, boxes, cyls, 0.05, False)[0][b]) / (2 * h)
        assert abs(fd - grad[b, j, w]) <= 1e-5 * max(1.0, abs(fd))


# ---- trajectory metrics: oracle/metrics_oracle.py against the reference's MetricsCalculator ----
def test_metrics_oracle_matches_reference(golden):
    from oracle import metrics_oracle as mo
    g = golden("metrics.npz")
    traj, dts = g["traj"], g["dts"]
    T = mo.ee_transforms(np.transpose(traj, (0, 2, 1)))
    np.testing.assert_allclose(T, g["ee_transforms"], rtol=0, atol=1e-6)
    for r in range(traj.shape[0]):
        np.testing.assert_allclose(mo.path_lengths(traj[r]), g["path_lengths"][r], rtol=1e-6, atol=1e-7)
        for i, dt in enumerate(dts):
            m = mo.trajectory_metrics(traj[r], float(dt))
            # SPARC is a float32 FFT inside the reference for the end-effector profile: 1e-5 relative
            np.testing.assert_allclose(m[2:], g["sparc"][i, r], rtol=1e-5, atol=1e-7)
    sal, f, Mf, first, last = mo.sparc(mo.speed_profiles(traj[0], float(dts[0]))[0], 1. / float(dts[0]))
    np.testing.assert_allclose(f, g["f"], rtol=0, atol=0)
    np.testing.assert_allclose(Mf, g["Mf_joint"], rtol=0, atol=1e-12)
    np.testing.assert_array_equal(f[first:last + 1], g["fsel_joint"])
    # the known-answer example in the reference's docstring (lib/metrics.py:79-84): '%.5f' % sal == '-1.41403'
    ex = mo.sparc(g["example_move"], 100.)[0]
    assert "%.5f" % ex == "-1.41403"
    np.testing.assert_allclose(ex, float(g["example_sal"]), rtol=1e-12)
    # all-zero movement -> 0 (lib/metrics.py:86-88)
    assert mo.sparc(np.zeros(49), 50.)[0] == 0.0


@pytest.mark.parametrize("case", ["iv", "mixed", "bench10", "normal", "uncond"])
def test_oracle_single_steps_match_reference_tapes(golden, case):
    """tests/golden/tape_<case>.npz holds all 255 steps of a reference chain (state in -> reference step -> state out,
    oracle/make_golden.py `tapes`); the oracle restatement reproduces a spread of them (every 12th step and the last
    five, guided and unguided, incl. t = 1 with its row-0 noise quirk) to 2e-6 rad relative to the state's size."""
    from oracle.make_golden import TAPE_CASES, tape_case_inputs
    g = golden("tape_%s.npz" % case)
    guides, bpg, scene, sd, x_T, noise, condition = tape_case_inputs(case)
    assert abs(sum(n.sum() for n in noise) - float(g["noise_checksum"])) < 1e-9
    cfgs = so.expand_guide_tables([guide_params.GUIDES[n] for n in guides], bpg)
    beta, alpha, abar = so.schedule()
    tape = g["tape"].astype(np.float64)
    assert np.array_equal(tape[0][:, :, 1:-1].astype(np.float32), x_T[:, :, 1:-1].astype(np.float32))
    for k in sorted(set(range(0, 255, 12)) | {250, 251, 252, 253, 254}):
        t = 255 - k
        X = tape[k]
        with torch.no_grad():
            eps = unet_oracle.unet_forward(sd, torch.tensor(X, dtype=torch.float32), t).numpy()
        X = so.posterior_step(X, t, eps, noise[k], beta, alpha, abar)
        if t % 2 == 0 and t >= 5:
            with np.errstate(all="ignore"):
                G = go.gradient_analytic(so.clip_joints(X[:, :, 1:-1]), scenes.START, scenes.GOAL, scene, cfgs, t)
            X[:, :, 1:-1] -= cfgs["guidance_schedule"][:, t - 1, None, None] * G
        if condition:
            X[:, :, 0], X[:, :, -1] = scenes.START, scenes.GOAL
        scale = 2e-6 + 2e-7 * np.abs(tape[k + 1])
        assert np.all(np.abs(X - tape[k + 1]) <= scale), "case %s step t=%d: %g" % (
            case, t, np.max(np.abs(X - tape[k + 1]) / scale))
