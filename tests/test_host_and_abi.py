"""CPU: host-side logic and the C-ABI surface (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from edmp_b200 import _lib
    header = open(os.path.join(ROOT, "include", "edmp_b200.h")).read()
    declared = set(re.findall(r"\b(edmp_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.edmp_version() >= 100
    dims = (ctypes.c_int * 6)(32, 64, 128, 256, 512, 512)
    assert lib.edmp_unet_param_count(dims, 6) == 29938471


def test_no_cpu_fallback():
    import torch
    from edmp_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU box")
    with pytest.raises(_lib.EdmpError):
        _lib.require_cuda("cpu")
    with pytest.raises(_lib.EdmpError):
        _lib.require_cuda("cuda:0")


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "edmp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_key_table_matches_oracle_layout():
    from edmp_b200.diffusion import unet_key_table
    from oracle import weights
    mine = unet_key_table()
    ref = [(k, s) for k, s, _ in weights.key_table()]
    assert mine == ref


def test_guide_tables_match_oracle_and_shipped_yaml():
    from edmp_b200 import build_guide_cfgs, load_guide_hparams
    from oracle import guide_params, sampler_oracle as so
    guides = sorted(guide_params.GUIDES)
    hp = load_guide_hparams(guides, os.path.join(ROOT, "guides") + "/")
    for n, h in zip(guides, hp):
        h = dict(h)
        h.pop("batch_size", None)
        ref = guide_params.GUIDES[n]
        assert h["guidance_method"] == ref["guidance_method"] and bool(h["grad_norm"]) == ref["grad_norm"]
        assert [float(v) for v in h["obstacle_clearance"]["range"]] == ref["obstacle_clearance"]["range"]
        for k, v in ref["obstacle_expansion"].items():
            assert [float(x) for x in h["obstacle_expansion"][k]] == [float(x) for x in v]
        assert h["guidance_schedule"]["type"] == ref["guidance_schedule"]["type"]
        assert float(h["guidance_schedule"]["scale_val"]) == ref["guidance_schedule"]["scale_val"]
    a = build_guide_cfgs(hp, 3)
    b = so.expand_guide_tables([guide_params.GUIDES[n] for n in guides], 3)
    for k in b:
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k


@pytest.mark.reference
def test_shipped_yaml_equals_reference_yaml():
    for f in os.listdir("/root/reference/guides/cfgs"):
        ref = yaml.safe_load(open(os.path.join("/root/reference/guides/cfgs", f)))
        mine = yaml.safe_load(open(os.path.join(ROOT, "guides", "cfgs", f)))
        assert ref == mine, f


def test_unet_checkpoint_roundtrip(tmp_path):
    import torch
    from edmp_b200 import TemporalUNet
    from oracle import weights
    d = str(tmp_path / "TemporalUNetModel255_N50")
    m = TemporalUNet(d, 7, 32, "cuda:0", dims=(32, 64, 128, 256, 512, 512))
    assert os.path.isdir(d) and m.losses.size == 0
    sd = weights.seeded_state_dict(3)
    m.load_state_dict(sd)
    m.losses = np.zeros(2)
    m.save()
    m2 = TemporalUNet(d, 7, 32, "cuda:0", dims=(32, 64, 128, 256, 512, 512))
    assert m2.losses.size == 2
    for k, v in m2.state_dict().items():
        assert torch.equal(v, sd[k])
    assert m2.flat_params().size == 29938471
    with pytest.raises(RuntimeError):
        bad = dict(sd)
        bad.pop("final_conv.1.bias")
        m2.load_state_dict(bad)


def test_synthetic_inputs_match_the_oracle_generators():
    """bench.py's GPU arm takes its inputs from edmp_b200.synthetic, the CPU arm (oracle) from oracle.scenes /
    oracle.weights: both arms must see bit-identical problems and weights."""
    import torch
    from edmp_b200 import synthetic as S
    from oracle import sampler_oracle as so, scenes, weights
    assert np.array_equal(S.synthetic_scene(20, 1, True, 4), scenes.synthetic_scene(20, seed=1, rotated=True, cylinders=4))
    assert np.array_equal(S.tabletop_scene(), scenes.tabletop_scene())
    assert np.array_equal(S.START, scenes.START) and np.array_equal(S.GOAL, scenes.GOAL)
    _, _, abar = so.schedule()
    assert S.alpha_bar_T() == abar[-1]
    assert np.array_equal(S.gentle_x_T(12, abar[-1], seed=100), scenes.gentle_x_T(12, abar[-1], seed=100))
    a, b = S.seeded_state_dict(0, final_gain=0.2), weights.seeded_state_dict(0, final_gain=0.2)
    assert list(a.keys()) == list(b.keys()) and all(torch.equal(a[k], b[k]) for k in a)


def test_infer_serial_cfg_surface():
    """The entry point's host-side plumbing: benchmark cfg -> guide tables, synthetic problem source."""
    import infer_serial
    from edmp_b200 import YamlConfig
    cfg = YamlConfig(os.path.join(ROOT, "benchmark", "cfgs", "cfg_c1_synthetic.yaml"))
    problems, scene_types = infer_serial.load_problems(cfg)
    scene, start, goals = problems.fetch_data(0, scene_types[0])
    assert scene.shape[1] == 10 and start.shape == (7,) and goals.shape[1] == 7
    shipped = YamlConfig(os.path.join(ROOT, "benchmark", "cfgs", "cfg1.yaml"))     # the reference's paper config
    assert shipped["guide"]["guides"] == [1, 2, 3, 4, 5, 10, 11, 13, 14, 16, 18, 21]
    with pytest.raises(SystemExit):
        infer_serial.load_problems(shipped)   # 'hybrid' needs the reference's loader + downloads


def test_infer_serial_walks_every_scene_by_default():
    """`num_scenes_per_type: -1` (the shipped cfg1.yaml) and a missing key mean ALL scenes, like the reference's loop
    over dataset.data_nums[scene_type] (infer_serial.py:98); a positive value caps the walk."""
    import infer_serial
    from edmp_b200 import YamlConfig

    class Problems:
        data_nums = {"tabletop": 7, "cubby": 3}

    shipped = YamlConfig(os.path.join(ROOT, "benchmark", "cfgs", "cfg1.yaml"))
    assert shipped["dataset"]["num_scenes_per_type"] == -1
    assert infer_serial.scenes_per_type(shipped, Problems, "tabletop") == 7
    assert infer_serial.scenes_per_type(shipped, Problems, "cubby") == 3
    cfg = {"dataset": {}}
    assert infer_serial.scenes_per_type(cfg, Problems, "tabletop") == 7
    cfg = {"dataset": {"num_scenes_per_type": 2}}
    assert infer_serial.scenes_per_type(cfg, Problems, "tabletop") == 2
    assert infer_serial.scenes_per_type(cfg, Problems, "cubby") == 2
    cfg = {"dataset": {"num_scenes_per_type": 100}}
    assert infer_serial.scenes_per_type(cfg, Problems, "cubby") == 3
    # the synthetic problem source has a documented size for every scene type
    syn = infer_serial.SyntheticProblems()
    assert infer_serial.scenes_per_type({"dataset": {"num_scenes_per_type": -1}}, syn, "tabletop") == \
        infer_serial.SYNTHETIC_SCENES_PER_TYPE > 0


def test_obstacle_flattening_like_the_reference_loader():
    """edmp_b200.scene.flatten_obstacles against a literal restatement of datasets/load_test_dataset.py:105-151 on the
    fields the reference reads off geometrout's Cuboid / Cylinder (center, wxyz quaternion, dims | radius, height)."""
    from edmp_b200 import scene
    rng = np.random.default_rng(0)
    cuboids = [(rng.normal(size=3), rng.normal(size=4), rng.uniform(0.1, 0.5, 3)) for _ in range(3)]
    cylinders = [(rng.normal(size=3), rng.normal(size=4), rng.uniform(0.05, 0.2), rng.uniform(0.1, 0.6)) for _ in range(2)]
    cfg, cub, cyl, nb, nc = scene.flatten_obstacles(cuboids, cylinders)
    assert (nb, nc) == (3, 2) and cfg.shape == (5, 10) and cub.shape == (3, 10) and cyl.shape == (2, 9)
    # the reference's steps, one by one
    c_cent = np.array([c for c, _, _ in cuboids]); c_dims = np.array([d for _, _, d in cuboids])
    c_quat = np.roll(np.array([q for _, q, _ in cuboids]), -1, axis=1)                       # :128
    y_cent = np.array([c for c, _, _, _ in cylinders])
    y_quat = np.roll(np.array([q for _, q, _, _ in cylinders]), -1, axis=1)                  # :135
    y_r = np.array([[r] for _, _, r, _ in cylinders]); y_h = np.array([[h] for _, _, _, h in cylinders])
    y_dims = np.concatenate([y_r, y_r, y_h], axis=1)                                         # :136-139
    want = np.concatenate([np.concatenate([c_cent, y_cent]), np.concatenate([c_quat, y_quat]),
                           np.concatenate([c_dims, y_dims])], axis=1)                        # :141-151
    np.testing.assert_array_equal(cfg, want)
    np.testing.assert_array_equal(cub, np.concatenate([c_cent, c_quat, c_dims], axis=1))     # :129
    np.testing.assert_array_equal(cyl, np.concatenate([y_cent, y_quat, y_r, y_h], axis=1))   # :136
    # cuboids only / cylinders only / empty
    only_b = scene.flatten_obstacles(cuboids, ())
    assert only_b[0].shape == (3, 10) and only_b[2].shape == (0, 9) and only_b[4] == 0
    only_c = scene.flatten_obstacles((), cylinders)
    assert only_c[0].shape == (2, 10) and only_c[3] == 0
    with pytest.raises(ValueError):
        scene.flatten_obstacles((), ())


def test_goal_selection_like_the_entry_point():
    """edmp_b200.scene.select_goal: infer_serial.py:122-129 (volume trust region, then joint-space distance)."""
    from edmp_b200 import scene
    start = np.zeros(7)
    goals = np.stack([np.full(7, 3.0), np.full(7, 1.0), np.full(7, 0.5), np.full(7, 2.0)])
    vol = np.array([0.0, 0.0005, 0.01, 0.0007999])
    goal, idx = scene.select_goal(vol, start, goals)
    assert idx == 1 and np.array_equal(goal, goals[1])        # 2 is nearest but outside the trust region
    goal, idx = scene.select_goal(vol, start, goals, trust_region=0.02)
    assert idx == 2
    goal, idx = scene.select_goal(np.array([0.3]), start, goals[:1])
    assert idx == 0
    with pytest.raises(ValueError):
        scene.select_goal(vol[:2], start, goals)


@pytest.mark.parametrize("line,ref_line", [("r1_final_8190rows_bench.json", "r1_final_reference_arm_bench.json"),
                                           ("r2_final_8190rows_bench.json", "r2_final_bench_reference.json")])
def test_committed_bench_line_has_the_contract_keys(line, ref_line):
    """The bench lines committed under profiles/ (written by bench.py on a B200) carry every key of the driver's contract."""
    import json
    d = json.load(open(os.path.join(ROOT, "profiles", line)))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["warmup"] >= 3 and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["value"] != d["value"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["bound"] in ("hbm", "tensor")
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0
    assert d["gpu_launches"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if line.startswith("r2"):
        # round 2: strong-scaling block, the Python-API end-to-end numbers, the stock PyTorch-CUDA baseline
        assert d["strong"]["rows_total"] > 0 and d["e2e_api"]["philox"] > 0 and d["e2e_api"]["numpy"] > 0
        assert d["gpu_baseline"]["fp32"] > 0 and d["roofline"]["frac"] <= 1.0 / 3.0 + 1e-9   # the x3 modes' cap
        d8 = json.load(open(os.path.join(ROOT, "profiles", "r2_8gpu_bench.json")))
        assert d8["n_gpus"] == 8 and d8["metric"] == d["metric"] and d8["strong"]["rows_per_gpu"] * 8 == d8["strong"]["rows_total"]
    ref = json.load(open(os.path.join(ROOT, "profiles", ref_line)))
    assert ref["impl"] == "reference" and ref["metric"] == d["metric"] and ref["unit"] == d["unit"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["cpu_baseline"]["value"] == ref["value"]


def test_metrics_nfft_matches_the_reference_formula():
    """edmp_metrics_nfft is host-only arithmetic: int(2 ** (ceil(log2(m)) + padlevel)) of lib/metrics.py:90."""
    from edmp_b200 import _lib
    lib = _lib.load()
    for m in (1, 2, 3, 48, 49, 63, 64, 65, 200, 1000):
        for pad in (0, 2, 4):
            assert lib.edmp_metrics_nfft(m, pad) == int(2 ** (np.ceil(np.log2(m)) + pad)), (m, pad)
    assert lib.edmp_metrics_nfft(0, 4) == 0 and lib.edmp_metrics_nfft(49, -1) == 0


def test_sparc_oracle_properties():
    """Size-independent properties of the spectral arc length (oracle side): invariant to the amplitude of the profile
    and to time reversal (the magnitude spectrum is), and smoother profiles score closer to zero."""
    from oracle import metrics_oracle as mo
    t = np.linspace(0, 1, 49)
    bell = np.exp(-((t - 0.5) / 0.18) ** 2)
    wobble = bell * (1 + 0.3 * np.sin(2 * np.pi * 6 * t))
    a = mo.sparc(bell, 50.)[0]
    np.testing.assert_allclose(mo.sparc(3.7 * bell, 50.)[0], a, rtol=1e-12)
    np.testing.assert_allclose(mo.sparc((bell + 0.2 * t)[::-1], 50.)[0], mo.sparc(bell + 0.2 * t, 50.)[0], rtol=1e-9)
    assert mo.sparc(wobble, 50.)[0] < a < 0


def test_obstacle_cap_is_an_error_not_a_truncation():
    """EDMP_MAX_OBSTACLES (64) is a hard cap of the scene tables: a larger scene is refused with a message (before any
    CUDA call, so this runs without a GPU), never silently truncated."""
    import ctypes
    from edmp_b200 import _lib
    lib = _lib.load()
    cfg = np.zeros((65, 10))
    cfg[:, 6] = 1.0
    cfg[:, 7:] = 0.1
    h = ctypes.c_void_p()
    rc = lib.edmp_scene_create(cfg.ctypes.data_as(ctypes.c_void_p), 65, None, ctypes.byref(h))
    assert rc != 0 and b"1..64" in lib.edmp_last_error()
    rc = lib.edmp_scene_create(cfg.ctypes.data_as(ctypes.c_void_p), 0, None, ctypes.byref(h))
    assert rc != 0
