"""TEST INFRASTRUCTURE ONLY.  Hyper-parameters of the shipped guide YAMLs, as data.

Values restate reference guides/cfgs/guide<N>.yaml (e.g. guide10.yaml:1-20) so oracle-side
tests do not depend on the product's YAML loader.  tests/test_guide_cfgs.py checks the product's
own guides/cfgs/*.yaml against this table (and, where /root/reference exists, this table
against the reference files).
"""


def _h(clr, e1, v1, e2, v2, e3, v3, method, gn, sched, scale):
    return {"obstacle_clearance": {"range": list(clr)},
            "obstacle_expansion": {"isr1": list(e1), "val1": list(v1), "isr2": list(e2),
                                   "val2": list(v2), "isr3": list(e3), "val3": list(v3)},
            "guidance_method": method, "grad_norm": gn,
            "guidance_schedule": {"type": sched, "scale_val": scale},
            "volume_trust_region": 0.0008}


_IV0 = ((150, 255), (0.0, 0.0), (20, 150), (0.0, 0.0), (0, 20), (0.0, 0.0))
_DEC = ((150, 255), (0.4, 0.4), (20, 150), (0.4, 0.0), (0, 20), (0.0, 0.0))
_INC = ((150, 255), (0.4, 0.4), (20, 150), (0.0, 0.4), (0, 20), (0.0, 0.0))
_LATE = ((40, 255), (0.4, 0.4), (10, 40), (0.0, 0.4), (0, 20), (0.0, 0.0))
GUIDES = {
    1: _h((0.1, 0.1), *_IV0, "iv", False, "varying", 0.05),
    2: _h((0.05, 0.05), *_IV0, "iv", False, "varying", 0.05),
    3: _h((0.01, 0.01), *_IV0, "iv", False, "varying", 0.05),
    4: _h((0.15, 0.15), *_IV0, "iv", False, "varying", 0.05),
    5: _h((0.01, 0.15), *_IV0, "iv", False, "varying", 0.05),
    9: _h((0.0, 0.0), *_DEC, "iv", True, "constant", 0.05),
    10: _h((0.06, 0.06), (80, 255), (0.4, 0.4), (20, 80), (0.0, 0.0), (0, 20), (0.0, 0.0),
           "sv", False, "varying", 0.05),
    11: _h((0.0, 0.0), *_INC, "sv", True, "constant", 0.05),
    12: _h((0.0, 0.0), *_DEC, "iv", True, "constant", 0.05),
    13: _h((0.0, 0.0), *_INC, "sv", True, "constant", 0.01),
    14: _h((0.02, 0.02), *_INC, "sv", True, "constant", 0.1),
    15: _h((0.0, 0.0), *_DEC, "iv", True, "constant", 0.05),
    16: _h((0.1, 0.1), *_INC, "sv", True, "constant", 0.1),
    17: _h((0.0, 0.0), *_DEC, "iv", True, "constant", 0.05),
    18: _h((0.05, 0.05), *_LATE, "sv", True, "constant", 0.05),
    21: _h((0.05, 0.05), *_LATE, "sv", True, "constant", 0.1),
}
