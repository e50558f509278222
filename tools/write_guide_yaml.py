"""Writes guides/cfgs/guide<N>.yaml (the plugin files of the reference's guides/cfgs/) from the
hyper-parameter table in oracle/guide_params.py.  `index` and `batch_size` are informational in the
reference (never read by infer_serial.py); they are kept, quirks included, so the parsed files are
identical to the reference's."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.guide_params import GUIDES  # noqa: E402

INDEX = {9: 1, 12: 1, 15: 1, 17: 1}
NO_BATCH_SIZE = {4}


def fmt(v):
    return "[%s]" % ", ".join(repr(float(x)) if isinstance(x, float) else repr(x) for x in v)


for n, h in sorted(GUIDES.items()):
    oe = h["obstacle_expansion"]
    lines = ["# guide %d -- cost hyper-parameters for one member of the guide ensemble" % n,
             "index: %d" % INDEX.get(n, n), "", "hyperparameters:"]
    if n not in NO_BATCH_SIZE:
        lines.append("  batch_size: 10")
    lines += ["  obstacle_clearance:",
              "    range: %s            # metres added to every obstacle extent, linear over t" % fmt(h["obstacle_clearance"]["range"]),
              "  obstacle_expansion:                # minimum obstacle extent per diffusion-step segment"]
    for k in ("1", "2", "3"):
        lines.append("    isr%s: %s" % (k, fmt(oe["isr" + k])))
        lines.append("    val%s: %s" % (k, fmt(oe["val" + k])))
    lines += ["  guidance_method: '%s'              # iv = intersection volume, sv = swept volume" % h["guidance_method"],
              "  grad_norm: %s" % ("True" if h["grad_norm"] else "False"),
              "  guidance_schedule:",
              "    type: '%s'                 # 'varying' = 1.4 + t/T, 'constant' = scale_val" % h["guidance_schedule"]["type"],
              "    scale_val: %r" % h["guidance_schedule"]["scale_val"],
              "  volume_trust_region: %r" % h["volume_trust_region"], ""]
    with open(os.path.join(ROOT, "guides", "cfgs", "guide%d.yaml" % n), "w") as f:
        f.write("\n".join(lines))
print("wrote", len(GUIDES), "guide files")
