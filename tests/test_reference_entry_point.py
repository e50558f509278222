"""Build-container test (needs /root/reference, no GPU): the UNMODIFIED reference entry point
(/root/reference/infer_serial.py, executed with runpy as __main__) runs over the swapped packages -- `lib` and
`diffusion` are edmp_b200.lib / edmp_b200.diffusion -- with the C ABI mocked at the ctypes layer, so that every call site
of the script (infer_serial.py:1-2 star imports, :44 Diffusion, :50 TemporalUNet, :112 IntersectionVolumeGuide, :119 cost,
:134-143 denoise_guided, :147 choose_best_trajectory, :100-101,:159-165 the environment) is shown to bind: names,
keyword arguments, shapes and dtypes reach libedmp_b200's entry points with the values the reference passes.
INTEGRATION.md section A describes exactly this package swap."""
import contextlib
import ctypes
import os
import runpy
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ENTRY = "/root/reference/infer_serial.py"
pytestmark = pytest.mark.reference


class FakeLib:
    """Stands in for ctypes.CDLL(libedmp_b200.so): records every call, hands out handles, zero-fills outputs."""

    def __init__(self):
        from edmp_b200 import _lib
        self.calls = []
        self.n_obs = {}
        self._next = 0x1000
        self._names = set(_lib.SIGNATURES)

    def _handle(self, ref):
        self._next += 0x10
        ref._obj.value = self._next
        return self._next

    def __getattr__(self, name):
        if name.startswith("_") or name not in self._names:
            raise AttributeError(name)

        def call(*args):
            self.calls.append((name, args))
            return getattr(self, "do_" + name, lambda *a: 0)(*args)
        return call

    # ---- behaviours the host code depends on -------------------------------------------------------------------
    def do_edmp_last_error(self):
        return b""

    def do_edmp_unet_param_count(self, dims, n):
        from edmp_b200.diffusion import unet_key_table
        return sum(int(np.prod(shape)) for _, shape in unet_key_table(7, tuple(dims[i] for i in range(n))))

    def do_edmp_unet_create(self, params, n_params, dims, n_dims, precision, max_rows, out):
        self._handle(out)
        return 0

    # (the checkpoint comes from disk, so the host class packs and caches the weight blob: edmp_unet_pack + blob read)
    do_edmp_unet_pack = do_edmp_unet_create

    def do_edmp_unet_blob_bytes(self, unet):
        return 64

    def do_edmp_unet_blob_layout_version(self):
        return 1

    def do_edmp_sampler_create(self, T, thresh, rows, out):
        self._handle(out)
        return 0

    def do_edmp_scene_create(self, cfg, n_obs, link_dims, out):
        self.n_obs[self._handle(out)] = n_obs
        return 0

    def do_edmp_guide_volumes(self, scene, q, start, goal, t, mode, rows, n, out, stream):
        ctypes.memset(out.value, 0, rows * (n + 1 if mode else n) * 9 * self.n_obs[scene.value] * 4)
        return 0

    def do_edmp_sample_guided(self, sampler, unet, scene, x, start, goal, noise, seed, rows, t_start, t_stop, cost, stream):
        if cost is not None and cost.value:
            ctypes.memset(cost.value, 0, rows * 4)
        return 0

    def do_edmp_guide_final_cost(self, scene, traj, start, goal, rows, cost, stream):
        vals = (ctypes.c_float * rows)(*[float(rows - r) for r in range(rows)])     # the LAST row is the cheapest
        ctypes.memmove(cost.value, vals, rows * 4)
        return 0

    def do_edmp_unet_range_status(self, unet, flag, stream):
        flag._obj.value = 0
        return 0

    def do_edmp_sampler_last_launches(self, sampler):
        return 0


class FakeProblems:
    """datasets.load_test_dataset.TestDataset stand-in (the MpiNets pickles are a download): same constructor,
    data_nums and 7-tuple fetch_data (load_test_dataset.py:16,:57-61,:189)."""
    instances = []

    def __init__(self, type="global", d_path="./datasets/"):
        self.type, self.d_path = type, d_path
        self.data_nums = {"tabletop": 2}
        self.fetched = []
        FakeProblems.instances.append(self)

    def fetch_data(self, scene_num, scene_type="tabletop"):
        from edmp_b200 import synthetic
        self.fetched.append((scene_num, scene_type))
        scene = synthetic.tabletop_scene(seed=3 + scene_num)
        cuboids = scene.copy()
        return scene, cuboids, np.zeros((0, 10)), scene.shape[0], 0, synthetic.START.copy(), \
            synthetic.goal_candidates(6, seed=7 + scene_num)


@pytest.fixture
def swapped(monkeypatch, tmp_path):
    import edmp_b200
    import edmp_b200.diffusion
    import edmp_b200.lib
    from edmp_b200 import _lib
    fake = FakeLib()
    monkeypatch.setattr(_lib, "load", lambda: fake)
    monkeypatch.setattr(_lib, "require_cuda", lambda device: torch.device("cpu"))
    monkeypatch.setattr(_lib, "stream_ptr", lambda: None)
    monkeypatch.setattr(torch.cuda, "device", lambda dev: contextlib.nullcontext())
    autolab = types.ModuleType("autolab_core")
    autolab.YamlConfig = edmp_b200.YamlConfig
    datasets = types.ModuleType("datasets")
    loader = types.ModuleType("datasets.load_test_dataset")
    loader.TestDataset = FakeProblems
    datasets.load_test_dataset = loader
    FakeProblems.instances = []
    for name, mod in (("lib", edmp_b200.lib), ("diffusion", edmp_b200.diffusion), ("autolab_core", autolab),
                      ("wandb", types.ModuleType("wandb")), ("datasets", datasets),
                      ("datasets.load_test_dataset", loader)):
        monkeypatch.setitem(sys.modules, name, mod)
    return fake


def test_unmodified_reference_entry_point_runs_over_the_swapped_packages(swapped, tmp_path, monkeypatch, capsys):
    from edmp_b200 import TemporalUNet, _lib
    fake = swapped
    # a model directory in the reference's layout (infer_serial.py:45-49: <model_dir>TemporalUNetModel<T>_N<traj_len>)
    model_dir = str(tmp_path / "models") + "/"
    os.mkdir(model_dir)
    TemporalUNet(model_dir + "TemporalUNetModel255_N50", 7, 32, "cuda:0", dims=(32, 64, 128, 256, 512, 512)).save()
    cfg = tmp_path / "cfg.yaml"
    cfg.write_text("""
guide:
  guides: [1, 10, 11]
  batch_size_per_guide: 2
  guide_path: '%s/guides/'
dataset:
  path: './datasets/'
  dataset_type: 'hybrid'
  scene_types: ['tabletop']
  num_scenes_per_type: -1
  random_scenes: False
  save_scene_indices: True
model:
  model_dir: '%s'
  device: 'cuda:0'
  T: 255
  traj_len: 50
  num_channels: 7
general:
  gui: False
  save_dir: './results/'
  wandb:
    enable_wandb: False
    run_num: 1
    project_name: 'x'
""" % (ROOT, model_dir))
    monkeypatch.setattr(sys, "argv", ["infer_serial.py", "-c", str(cfg)])
    np.random.seed(0)
    runpy.run_path(REF_ENTRY, run_name="__main__")          # the reference's file, byte for byte

    out = capsys.readouterr().out
    assert out.count("Success: ") == 2 and "Denoiser time" in out and "IK time" in out
    assert FakeProblems.instances[0].type == "hybrid" and FakeProblems.instances[0].fetched == [(0, "tabletop"), (1, "tabletop")]
    names = [n for n, _ in fake.calls]
    by = {n: [a for m, a in fake.calls if m == n] for n in set(names)}
    total = 3 * 2
    # :50 TemporalUNet(model_name=, input_dim=, time_dim=32, dims=, device=): the engine is built once, with the whole
    # checkpoint, in the shipped default precision (the reference passes none) = the benchmarked f16x3 mode
    (params, n_params, dims, n_dims, precision, max_rows, _), = by["edmp_unet_pack"]
    assert len(os.listdir(os.path.join(model_dir, "TemporalUNetModel255_N50", "edmp_cache"))) == 1
    assert n_params == 29938471 and [dims[i] for i in range(n_dims)] == [32, 64, 128, 256, 512, 512]
    assert precision == _lib.PRECISIONS["f16x3"] and max_rows >= total
    # :112 IntersectionVolumeGuide(obstacle_config=, device=, guide_cfgs=, batch_size=): one scene per problem
    assert len(by["edmp_scene_create"]) == 2 and all(a[1] == 5 for a in by["edmp_scene_create"])
    # :119 guide.cost(tensor[k,7,1], 0, batch_size=k): t = 0, iv mode, k = 6 goal candidates, one waypoint
    for a in by["edmp_guide_volumes"]:
        assert (a[4], a[5], a[6], a[7]) == (0, 0, 6, 1)
    # :59-91 tables built by the reference's own loop reach the device for the whole flat batch as ONE ensemble
    for a in by["edmp_scene_set_guide_tables"]:
        assert (a[6], a[7]) == (total, total)
    # :134-143 denoise_guided(model=, guide=, batch_size=, traj_len=, num_channels=, condition=True, benchmarking=True,
    #                         start=, goal=, guidance_schedule=): all 255 steps in one call, recorded numpy noise tape
    #   (the numpy noise stream reaches the device in 16-step chunks: 16 calls per problem, contiguous in t)
    calls = by["edmp_sample_guided"]
    assert len(calls) == 2 * 16
    for prob in (calls[:16], calls[16:]):
        t = 255
        for a in prob:
            assert a[8] == total and a[9] == t and a[6] is not None          # rows, t_start, the noise chunk
            assert (a[11] is not None) == (a[10] == 0)                       # final costs with the last chunk only
            t = a[10]
        assert t == 0
    assert all(a[1] == 1 for a in by["edmp_sampler_set_condition"])
    # :147 choose_best_trajectory(start_joints, goal_joints, trajectories): argmin of the device-side final costs
    assert len(by["edmp_guide_final_cost"]) == 2 and all(a[4] == total for a in by["edmp_guide_final_cost"])
    # Diffusion(T=, device=) at :44 creates the sampler lazily with T = 255 for the batch
    assert by["edmp_sampler_create"][0][0] == 255 and by["edmp_sampler_create"][0][2] >= total
