"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (edmp_b200/).

Loads the *unmodified* reference implementation from /root/reference under
import shims so it can run in this container.  Nothing is
copied from the reference: its modules are imported from where they lie.

Shims (all harness side, none touch the reference files):
  1. stub ``matplotlib`` / ``matplotlib.pyplot`` (diffusion/gaussian.py:2).
  2. stub ``pybullet_data.getDataPath()`` -> temp dir holding
     franka_panda/meshes/collision/{link1..7,hand,finger}.obj, symlinked to
     robofin's hd_meshes (lib/guide.py:245); lib/guide.py is loaded by file path
     so lib/__init__.py (which imports pybullet) is bypassed.
  3. numpy-1.x semantics for ``np.where(<python bool>)`` in
     diffusion/diffusion.py:127 (numpy >= 2.1 raises on 0-d input).
  4. TemporalUNet(model_name=<writable tmp dir with weights_latest.pt+losses.npy>).
  5. record / replay of ``np.random.multivariate_normal`` so another
     implementation can consume the identical x_T and z_t.

This module only works where /root/reference exists (the build container); the
GPU box uses the committed fixtures under tests/golden/ instead.
"""
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np

REF_ROOT = os.environ.get("EDMP_REFERENCE_ROOT", "/root/reference")
_LINKS = ["link1", "link2", "link3", "link4", "link5", "link6", "link7", "hand", "finger"]
_state = {}


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "diffusion"))


def _install_stubs():
    if "stubs" in _state:
        return
    # (1) matplotlib
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
    # (2) pybullet_data -> temp mesh dir
    tmp = tempfile.mkdtemp(prefix="edmp_pbdata_")
    mesh_dst = os.path.join(tmp, "franka_panda", "meshes", "collision")
    os.makedirs(mesh_dst)
    mesh_src = os.path.join(REF_ROOT, "robofin", "robofin", "urdf", "franka_panda",
                            "hd_meshes", "collision")
    for name in _LINKS:
        os.symlink(os.path.join(mesh_src, name + ".obj"), os.path.join(mesh_dst, name + ".obj"))
    pbd = types.ModuleType("pybullet_data")
    pbd.getDataPath = lambda: tmp
    sys.modules["pybullet_data"] = pbd
    _state["stubs"] = tmp


class _NumpyCompat:
    """Proxy for the name ``np`` inside the reference's diffusion module: numpy 1.x
    ``where`` semantics for scalar conditions + optional noise record/replay."""

    def __init__(self, rng_hook):
        self._hook = rng_hook
        self.random = _RandomProxy(rng_hook)

    def where(self, cond, *a):
        if not a:
            return np.atleast_1d(np.asarray(cond)).nonzero()
        return np.where(cond, *a)

    def __getattr__(self, name):
        return getattr(np, name)


class _RandomProxy:
    def __init__(self, hook):
        self._hook = hook

    def multivariate_normal(self, mean, cov, size):
        return self._hook(mean, cov, size)

    def __getattr__(self, name):
        return getattr(np.random, name)


class NoiseTape:
    """Records the draws of np.random.multivariate_normal made by the reference's
    sampler (first draw = x_T, then one z per step t = T..1), or replays a tape."""

    def __init__(self, replay=None):
        self.replay = None if replay is None else [np.asarray(a, dtype=np.float64) for a in replay]
        self.draws = []
        self._i = 0

    def __call__(self, mean, cov, size):
        if self.replay is not None:
            out = self.replay[self._i].copy()
            self._i += 1
        else:
            out = np.random.multivariate_normal(mean=mean, cov=cov, size=size)
        self.draws.append(out.copy())
        return out


def load_reference():
    """Returns a namespace with the reference's Diffusion, TemporalUNet,
    IntersectionVolumeGuide classes and the modules they live in."""
    if "ns" in _state:
        return _state["ns"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import diffusion as ref_diffusion  # the reference package  (diffusion/__init__.py)
    spec = importlib.util.spec_from_file_location("_edmp_ref_guide",
                                                  os.path.join(REF_ROOT, "lib", "guide.py"))
    ref_guide = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_guide)
    ns = types.SimpleNamespace(
        Diffusion=ref_diffusion.Diffusion,
        TemporalUNet=ref_diffusion.TemporalUNet,
        IntersectionVolumeGuide=ref_guide.IntersectionVolumeGuide,
        diffusion_module=sys.modules["diffusion.diffusion"],
        guide_module=ref_guide,
    )
    _state["ns"] = ns
    return ns


def set_noise_tape(tape):
    """Routes the reference sampler's RNG draws through ``tape`` (and installs the
    numpy-1.x ``where`` shim).  Pass a fresh NoiseTape() to record."""
    ns = load_reference()
    ns.diffusion_module.np = _NumpyCompat(tape)
    return tape


def make_reference_unet(state_dict=None, device="cpu", seed=0):
    """Instantiates the reference TemporalUNet in a writable temp model dir.  With
    ``state_dict`` given it is loaded through the reference's own load() path
    (weights_latest.pt + losses.npy, temporalunet.py:88-92)."""
    import torch
    ns = load_reference()
    tmp = tempfile.mkdtemp(prefix="edmp_model_")
    model_dir = os.path.join(tmp, "TemporalUNetModel255_N50")
    if state_dict is not None:
        os.mkdir(model_dir)
        torch.save(state_dict, os.path.join(model_dir, "weights_latest.pt"))
        np.save(os.path.join(model_dir, "losses.npy"), np.zeros(1))
    else:
        torch.manual_seed(seed)
    model = ns.TemporalUNet(model_name=model_dir, input_dim=7, time_dim=32, device=device,
                            dims=(32, 64, 128, 256, 512, 512))
    model.train(False)
    return model


def reference_link_dimensions():
    """[9,3] link box table exactly as the reference derives it from the (stand-in)
    OBJ meshes (lib/guide.py:243-281)."""
    ns = load_reference()
    g = ns.IntersectionVolumeGuide(obstacle_config=np.zeros((1, 10)), device="cpu",
                                   guide_cfgs={}, batch_size=1)
    return g.link_dimensions.numpy().copy()
