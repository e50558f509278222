"""TEST INFRASTRUCTURE ONLY.  torch-fp32 CPU restatement of TemporalUNet.forward.

Follows reference diffusion/models/temporalunet.py:47-76 and blocks.py:13-34 (Conv1dBlock),
:38-54 (SinusoidalPosEmb), :58-72 (TimeMLP), :76-92 (TimeEmbedding), :137-166
(ResidualConvolutionBlock), :202-260 (Down/Middle/UpSampler), written functionally over a
state_dict.  Pinned against the reference module by tests/test_oracle_vs_reference.py and
the committed fixtures tests/golden/unet_*.npz (made by oracle/make_golden.py).
"""
import math

import torch
import torch.nn.functional as F


def time_embedding(sd, t):
    """t: float tensor [1] -> [1, 32]   (blocks.py:44-54, :83-88)"""
    half = 16
    scale = math.log(10000) / (half - 1)
    freqs = torch.exp(torch.arange(half) * -scale).to(t.device)
    ang = t[:, None] * freqs[None, :]
    emb = torch.cat((ang.sin(), ang.cos()), dim=-1)
    h = F.linear(emb, sd["time_embedding.time_mlp.1.weight"], sd["time_embedding.time_mlp.1.bias"])
    h = F.mish(h)
    return F.linear(h, sd["time_embedding.time_mlp.3.weight"], sd["time_embedding.time_mlp.3.bias"])


def _conv_block(sd, p, x):
    y = F.conv1d(x, sd[p + ".block.0.weight"], sd[p + ".block.0.bias"], padding=2)
    y = F.group_norm(y, 8, sd[p + ".block.2.weight"], sd[p + ".block.2.bias"], eps=1e-5)
    return F.mish(y)


def _res_block(sd, p, x, temb, taps=None):
    out = _conv_block(sd, p + ".blocks.0", x)
    tm = F.linear(F.mish(temb), sd[p + ".time_mlp.time_mlp.1.weight"], sd[p + ".time_mlp.time_mlp.1.bias"])
    out = out + tm[:, :, None]
    if taps is not None:
        taps[p + ".blocks.0"] = out
    out = _conv_block(sd, p + ".blocks.1", out)
    if (p + ".residual_conv.weight") in sd:
        res = F.conv1d(x, sd[p + ".residual_conv.weight"], sd[p + ".residual_conv.bias"])
    else:
        res = x
    out = out + res
    if taps is not None:
        taps[p] = out
    return out


def unet_forward(sd, x, t, taps=None):
    """x: [B,7,50] float32, t: python number or [1] tensor -> eps [B,7,50].
    ``taps`` (dict) optionally receives every block output for per-layer parity."""
    if not torch.is_tensor(t):
        t = torch.tensor([float(t)], dtype=torch.float32)
    # (device-agnostic: bench.py's secondary baseline runs this same restatement on cuda with stock PyTorch kernels)
    temb = time_embedding(sd, t.to(x.device, torch.float32))
    n_down = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("down_samplers."))
    n_up = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("up_samplers."))
    skips = []
    for i in range(n_down):
        x = _res_block(sd, "down_samplers.%d.down.0" % i, x, temb, taps)
        x = _res_block(sd, "down_samplers.%d.down.1" % i, x, temb, taps)
        skips.append(x)
        wk = "down_samplers.%d.down.3.weight" % i
        if wk in sd:
            x = F.conv1d(x, sd[wk], sd["down_samplers.%d.down.3.bias" % i], stride=2, padding=1)
            if taps is not None:
                taps["down_samplers.%d.down.3" % i] = x
    x = _res_block(sd, "middle_block.middle.0", x, temb, taps)
    x = _res_block(sd, "middle_block.middle.2", x, temb, taps)
    for i in range(n_up):
        x = torch.cat([x, skips.pop()], dim=1)
        x = _res_block(sd, "up_samplers.%d.up.0" % i, x, temb, taps)
        x = _res_block(sd, "up_samplers.%d.up.1" % i, x, temb, taps)
        x = F.conv_transpose1d(x, sd["up_samplers.%d.up.3.weight" % i],
                               sd["up_samplers.%d.up.3.bias" % i], stride=2, padding=1)
        if x.shape[2] in (8, 14, 26):        # temporalunet.py:70-71: drop the last column
            x = x[:, :, :-1]
        if taps is not None:
            taps["up_samplers.%d.up.3" % i] = x
    x = _conv_block(sd, "final_conv.0", x)
    if taps is not None:
        taps["final_conv.0"] = x
    return F.conv1d(x, sd["final_conv.1.weight"], sd["final_conv.1.bias"])
