"""TEST INFRASTRUCTURE ONLY.  Deterministic synthetic TemporalUNet weights.

The released EDMP weights are a Google-Drive download (reference README.md:53-58) and are
absent, so parity runs use a seeded state_dict whose keys/shapes follow the reference
module tree (diffusion/models/temporalunet.py:11-36, blocks.py:137-260).  Values are drawn
with torch's CPU generator so this container and the GPU box produce identical bits.
"""
import math
from collections import OrderedDict

import torch

DIMS = (32, 64, 128, 256, 512, 512)
TIME_DIM = 32


def _res_block(prefix, cin, cout, out):
    out.append((prefix + ".blocks.0.block.0.weight", (cout, cin, 5), "w"))
    out.append((prefix + ".blocks.0.block.0.bias", (cout,), "b%d" % (cin * 5)))
    out.append((prefix + ".blocks.0.block.2.weight", (cout,), "gamma"))
    out.append((prefix + ".blocks.0.block.2.bias", (cout,), "beta"))
    out.append((prefix + ".blocks.1.block.0.weight", (cout, cout, 5), "w"))
    out.append((prefix + ".blocks.1.block.0.bias", (cout,), "b%d" % (cout * 5)))
    out.append((prefix + ".blocks.1.block.2.weight", (cout,), "gamma"))
    out.append((prefix + ".blocks.1.block.2.bias", (cout,), "beta"))
    out.append((prefix + ".time_mlp.time_mlp.1.weight", (cout, TIME_DIM), "w"))
    out.append((prefix + ".time_mlp.time_mlp.1.bias", (cout,), "b%d" % TIME_DIM))
    if cin != cout:
        out.append((prefix + ".residual_conv.weight", (cout, cin, 1), "w"))
        out.append((prefix + ".residual_conv.bias", (cout,), "b%d" % cin))


def key_table(input_dim=7, dims=DIMS):
    """[(key, shape, kind)] in the reference's state_dict order."""
    d = [input_dim, *dims]
    out = [("time_embedding.time_mlp.1.weight", (4 * TIME_DIM, TIME_DIM), "w"),
           ("time_embedding.time_mlp.1.bias", (4 * TIME_DIM,), "b%d" % TIME_DIM),
           ("time_embedding.time_mlp.3.weight", (TIME_DIM, 4 * TIME_DIM), "w"),
           ("time_embedding.time_mlp.3.bias", (TIME_DIM,), "b%d" % (4 * TIME_DIM))]
    n_down = len(d) - 1
    for i in range(n_down):
        cin, cout = d[i], d[i + 1]
        _res_block("down_samplers.%d.down.0" % i, cin, cout, out)
        _res_block("down_samplers.%d.down.1" % i, cout, cout, out)
        if i != n_down - 1:
            out.append(("down_samplers.%d.down.3.weight" % i, (cout, cout, 3), "w"))
            out.append(("down_samplers.%d.down.3.bias" % i, (cout,), "b%d" % (cout * 3)))
    mid = d[-1]
    _res_block("middle_block.middle.0", mid, mid, out)
    _res_block("middle_block.middle.2", mid, mid, out)
    for n, i in enumerate(range(len(d) - 1, 1, -1)):
        dim_in, dim_out = d[i - 1], d[i]
        _res_block("up_samplers.%d.up.0" % n, 2 * dim_out, dim_in, out)
        _res_block("up_samplers.%d.up.1" % n, dim_in, dim_in, out)
        # ConvTranspose1d weight is [C_in, C_out, k]; torch's fan_in for it is C_out*k
        out.append(("up_samplers.%d.up.3.weight" % n, (dim_in, dim_in, 4), "w"))
        out.append(("up_samplers.%d.up.3.bias" % n, (dim_in,), "b%d" % (dim_in * 4)))
    out.append(("final_conv.0.block.0.weight", (d[1], d[1], 5), "w"))
    out.append(("final_conv.0.block.0.bias", (d[1],), "b%d" % (d[1] * 5)))
    out.append(("final_conv.0.block.2.weight", (d[1],), "gamma"))
    out.append(("final_conv.0.block.2.bias", (d[1],), "beta"))
    out.append(("final_conv.1.weight", (input_dim, d[1], 1), "w"))
    out.append(("final_conv.1.bias", (input_dim,), "b%d" % d[1]))
    return out


def seeded_state_dict(seed=0, gain=1.0, input_dim=7, dims=DIMS, final_gain=1.0):
    """U(-1/sqrt(fan_in), 1/sqrt(fan_in)) * gain for conv/linear weights and biases (the
    torch default-init bound), GroupNorm gamma = 1 + 0.2*U(-1,1), beta = 0.2*U(-1,1) so the
    affine path is exercised.  ``final_gain`` scales the last 1x1 conv: an untrained net's eps is
    an arbitrary drift, and 0.2 keeps the 255-step chain near the joint limits' interior so the
    end-to-end fixtures are well conditioned ("conditioning of the chain")."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    sd = OrderedDict()
    for key, shape, kind in key_table(input_dim, dims):
        u = torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0
        if kind == "w":
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            sd[key] = u * (gain / math.sqrt(fan_in))
        elif kind[0] == "b" and kind[1:].isdigit():
            sd[key] = u * (1.0 / math.sqrt(int(kind[1:])))
        elif kind == "gamma":
            sd[key] = 1.0 + 0.2 * u
        else:
            sd[key] = 0.2 * u
    if final_gain != 1.0:
        sd["final_conv.1.weight"] = sd["final_conv.1.weight"] * final_gain
        sd["final_conv.1.bias"] = sd["final_conv.1.bias"] * final_gain
    return sd
