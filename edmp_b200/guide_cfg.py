"""Guide / benchmark configuration plumbing (host side).

Mirrors the reference's plugin surface: benchmark/cfgs/cfg*.yaml + guides/cfgs/guide<N>.yaml read
through ``YamlConfig`` (autolab_core in the reference, infer_serial.py:7,25,73), the ``Guide``
wrapper (guides/guide_cfg.py:27-29) and the per-row table expansion of infer_serial.py:56-91.
"""
import os

import numpy as np
import yaml


class YamlConfig(dict):
    """Minimal stand-in for autolab_core.YamlConfig: a dict loaded from a YAML file (only item
    access is used by the reference entry point)."""

    def __init__(self, filename=None):
        super().__init__()
        self.filename = filename
        if filename is not None:
            with open(filename, "r") as f:
                self.update(yaml.safe_load(f) or {})

    @property
    def config(self):
        return self


class Guide:
    def __init__(self, path_to_yaml="./guides/cfgs/guide1.yaml") -> None:
        self.cfg = YamlConfig(path_to_yaml)


def load_guide_hparams(guides, guide_path="./guides/"):
    """``hyperparameters`` dicts of guide<N>.yaml for the listed guide indices."""
    out = []
    for g in guides:
        path = os.path.join(guide_path, "cfgs", "guide%s.yaml" % g)
        out.append(YamlConfig(path)["hyperparameters"])
    return out


def build_guide_cfgs(guide_hparams, batch_size_per_guide, T=255):
    """Per-row tables exactly as infer_serial.py:56-91 builds them: row r belongs to guide
    r // batch_size_per_guide; expansion segments are written isr1 -> isr2 -> isr3 so later ones
    win on overlap (guide18/21); 'varying' schedule = 1.4 + arange(T)/T indexed by t-1."""
    n_guides = len(guide_hparams)
    bpg = int(batch_size_per_guide)
    total = int(n_guides * bpg)
    cfgs = {"batch_size_per_guide": bpg,
            "total_batch_size": total,
            "clearance": np.zeros((total, T)),
            "expansion": np.zeros((total, T)),
            "guidance_method": np.zeros((total,)),
            "grad_norm": np.zeros((total,)),
            "guidance_schedule": np.zeros((total, T)),
            "volume_trust_region": np.zeros((total,))}
    for i, hp in enumerate(guide_hparams):
        lo, hi = i * bpg, (i + 1) * bpg
        c0, c1 = hp["obstacle_clearance"]["range"]
        cfgs["clearance"][lo:hi, :] = np.linspace(c0, c1, T)
        oe = hp["obstacle_expansion"]
        for seg in ("1", "2", "3"):
            a, b = oe["isr" + seg]
            v0, v1 = oe["val" + seg]
            cfgs["expansion"][lo:hi, a:b] = np.linspace(v0, v1, num=abs(b - a))
        cfgs["guidance_method"][lo:hi] = 1 if hp["guidance_method"] == "sv" else 0
        cfgs["grad_norm"][lo:hi] = 1 if hp["grad_norm"] else 0
        sched = hp["guidance_schedule"]
        cfgs["guidance_schedule"][lo:hi, :] = (1.4 + np.arange(T) / T) if sched["type"] == "varying" \
            else sched["scale_val"]
        cfgs["volume_trust_region"][lo:hi] = hp["volume_trust_region"]
    return cfgs
