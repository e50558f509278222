"""TemporalUNet -- reference-shaped boundary class over the CUDA engine.

Same constructor, ``forward(x, t)``, checkpoint file layout and state_dict keys as reference
diffusion/models/temporalunet.py:9-100, but it is not an nn.Module: the parameters live in a plain
ordered dict and every forward runs in libedmp_b200.so (edmp_unet_forward).  There is no CPU path.
"""
import ctypes
import hashlib
import math
import os
import warnings
from collections import OrderedDict

import numpy as np
import torch

from .. import _lib

TIME_DIM = 32


def _res_block_keys(prefix, cin, cout, out):
    for blk, c_in in ((0, cin), (1, cout)):
        out.append(("%s.blocks.%d.block.0.weight" % (prefix, blk), (cout, c_in, 5)))
        out.append(("%s.blocks.%d.block.0.bias" % (prefix, blk), (cout,)))
        out.append(("%s.blocks.%d.block.2.weight" % (prefix, blk), (cout,)))
        out.append(("%s.blocks.%d.block.2.bias" % (prefix, blk), (cout,)))
    out.append((prefix + ".time_mlp.time_mlp.1.weight", (cout, TIME_DIM)))
    out.append((prefix + ".time_mlp.time_mlp.1.bias", (cout,)))
    if cin != cout:
        out.append((prefix + ".residual_conv.weight", (cout, cin, 1)))
        out.append((prefix + ".residual_conv.bias", (cout,)))


def unet_key_table(input_dim=7, dims=(32, 64, 128, 256, 512, 512)):
    """[(state_dict key, shape)] in checkpoint order (temporalunet.py:11-36 module tree)."""
    d = [input_dim, *dims]
    out = [("time_embedding.time_mlp.1.weight", (4 * TIME_DIM, TIME_DIM)),
           ("time_embedding.time_mlp.1.bias", (4 * TIME_DIM,)),
           ("time_embedding.time_mlp.3.weight", (TIME_DIM, 4 * TIME_DIM)),
           ("time_embedding.time_mlp.3.bias", (TIME_DIM,))]
    last = len(d) - 2
    for i in range(len(d) - 1):
        _res_block_keys("down_samplers.%d.down.0" % i, d[i], d[i + 1], out)
        _res_block_keys("down_samplers.%d.down.1" % i, d[i + 1], d[i + 1], out)
        if i != last:
            out.append(("down_samplers.%d.down.3.weight" % i, (d[i + 1], d[i + 1], 3)))
            out.append(("down_samplers.%d.down.3.bias" % i, (d[i + 1],)))
    _res_block_keys("middle_block.middle.0", d[-1], d[-1], out)
    _res_block_keys("middle_block.middle.2", d[-1], d[-1], out)
    for n, i in enumerate(range(len(d) - 1, 1, -1)):
        _res_block_keys("up_samplers.%d.up.0" % n, 2 * d[i], d[i - 1], out)
        _res_block_keys("up_samplers.%d.up.1" % n, d[i - 1], d[i - 1], out)
        out.append(("up_samplers.%d.up.3.weight" % n, (d[i - 1], d[i - 1], 4)))
        out.append(("up_samplers.%d.up.3.bias" % n, (d[i - 1],)))
    out.append(("final_conv.0.block.0.weight", (d[1], d[1], 5)))
    out.append(("final_conv.0.block.0.bias", (d[1],)))
    out.append(("final_conv.0.block.2.weight", (d[1],)))
    out.append(("final_conv.0.block.2.bias", (d[1],)))
    out.append(("final_conv.1.weight", (input_dim, d[1], 1)))
    out.append(("final_conv.1.bias", (input_dim,)))
    return out


def _default_init(table, generator=None):
    """torch's default Conv/Linear init bound (U(+-1/sqrt(fan_in))) and GroupNorm (1, 0) so a fresh
    model directory behaves like the reference's untrained module."""
    sd = OrderedDict()
    fan_in = 1
    for key, shape in table:
        if key.endswith(".block.2.weight"):
            sd[key] = torch.ones(shape)
        elif key.endswith(".block.2.bias"):
            sd[key] = torch.zeros(shape)
        else:
            if len(shape) > 1:
                fan_in = int(np.prod(shape[1:]))
            bound = 1.0 / math.sqrt(fan_in)
            sd[key] = (torch.rand(shape, generator=generator) * 2 - 1) * bound
    return sd


class TemporalUNet:

    #: environment knobs that change which kernels / tile shapes the engine plans (and therefore the packed layouts)
    PLAN_ENV = ("EDMP_CG2", "EDMP_NO_CG2", "EDMP_NO_NARROW", "EDMP_TC_V1", "EDMP_PM_V1", "EDMP_NO_PM")

    def __init__(self, model_name, input_dim, time_dim, device, dims=(32, 64, 128, 256),
                 precision=None, max_rows=64, blob_cache=None):
        """``precision``: arithmetic mode of the contractions.  The reference's call site passes none
        (infer_serial.py:50), so the default is the benchmarked parity-grade tensor-core mode: ``f16x3`` (tcgen05,
        IEEE-half hi/lo operand split), or whatever ``EDMP_PRECISION`` names (``fp32`` = CUDA-core kernels)."""
        if time_dim != TIME_DIM:
            raise ValueError("time_dim must be 32 (infer_serial.py:50)")
        if precision is None:
            precision = os.environ.get("EDMP_PRECISION", "f16x3")
        if precision not in _lib.PRECISIONS:
            raise ValueError("unknown precision %r (one of %s)" % (precision, sorted(_lib.PRECISIONS)))
        self.input_dim, self.time_dim, self.dims = int(input_dim), int(time_dim), tuple(int(d) for d in dims)
        self.device = device
        self.precision = precision
        self._max_rows = int(max_rows)
        self._handle = None
        #: versioned on-disk cache of the packed-weight blob under <model_name>/edmp_cache.
        #: None = on for weights that came from the checkpoint on disk (load()), off for in-memory state_dicts;
        #: EDMP_BLOB_CACHE=0 / 1 overrides.
        self.blob_cache = blob_cache
        self._from_checkpoint = False
        self._sha = None
        self.engine_source = None      # "state_dict" | "blob" | "packed" (how the current engine was built)
        self._table = unet_key_table(self.input_dim, self.dims)
        self.training = False
        self.model_name = model_name
        if not os.path.exists(model_name):
            os.mkdir(model_name)
            self.losses = np.array([])
            self._sd = _default_init(self._table)
        else:
            self._sd = None
            self.load()

    # ---- nn.Module-shaped surface used by the reference callers -----------------------------------
    def train(self, mode=True):
        self.training = bool(mode)
        return self

    def eval(self):
        return self.train(False)

    def to(self, device):
        self.device = device
        return self

    def state_dict(self):
        return OrderedDict((k, v.clone()) for k, v in self._sd.items())

    def load_state_dict(self, sd):
        keys = [k for k, _ in self._table]
        missing = [k for k in keys if k not in sd]
        unexpected = [k for k in sd if k not in set(keys)]
        if missing or unexpected:
            raise RuntimeError("state_dict mismatch: missing %s unexpected %s" % (missing[:3], unexpected[:3]))
        new = OrderedDict()
        for k, shape in self._table:
            v = torch.as_tensor(sd[k]).detach().to("cpu", torch.float32)
            if tuple(v.shape) != tuple(shape):
                raise RuntimeError("size mismatch for %s: %s vs %s" % (k, tuple(v.shape), shape))
            new[k] = v.contiguous()
        self._sd = new
        self._from_checkpoint = False
        self._sha = None
        self._release()

    def parameters(self):
        return iter(self._sd.values())

    def save(self):
        torch.save(self.state_dict(), self.model_name + "/weights_latest.pt")
        np.save(self.model_name + "/losses.npy", self.losses)

    def save_checkpoint(self, checkpoint):
        torch.save(self.state_dict(), self.model_name + "/weights_" + str(checkpoint) + ".pt")
        np.save(self.model_name + "/latest_checkpoint.npy", checkpoint)

    def load(self):
        self.losses = np.load(self.model_name + "/losses.npy")
        self.load_state_dict(torch.load(self.model_name + "/weights_latest.pt", map_location="cpu"))
        self._from_checkpoint = True
        print("Loaded Model at " + str(self.losses.size) + " epochs")

    def load_checkpoint(self, checkpoint):
        self.load_state_dict(torch.load(self.model_name + "/weights_" + str(checkpoint) + ".pt",
                                        map_location="cpu"))
        self.losses = np.load(self.model_name + "/losses.npy")[:checkpoint]

    # ---- engine -------------------------------------------------------------------------------------
    def _release(self):
        if self._handle is not None:
            _lib.load().edmp_unet_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def flat_params(self):
        """All tensors flattened and concatenated in state_dict order: what edmp_unet_create takes."""
        return torch.cat([v.reshape(-1) for v in self._sd.values()]).contiguous().numpy()

    # ---- packed-weight blob cache ------------------------------------------------------
    def _cache_enabled(self):
        env = os.environ.get("EDMP_BLOB_CACHE")
        if env is not None:
            return env not in ("0", "")
        return self._from_checkpoint if self.blob_cache is None else bool(self.blob_cache)

    @staticmethod
    def plan_class(max_rows):
        """The batch-size class the engine plans its tiles for: narrow column tiles below ~1300 rows (one class per
        row-tile count up to 40), CTA pairs from 4096 rows on, plus the planning knobs of the environment."""
        rts = (int(max_rows) + 127) // 128
        env = "".join("%s=%s;" % (k, os.environ[k]) for k in TemporalUNet.PLAN_ENV if k in os.environ)
        tag = "r%d%s" % (min(rts, 40), "p" if max_rows >= 4096 else "")
        return tag + ("-" + hashlib.sha256(env.encode()).hexdigest()[:8] if env else "")

    def blob_path(self, max_rows, version=None):
        """<model dir>/edmp_cache/unet_<sha256(state_dict)[:16]>_<precision>_v<layout version>_<plan class>.blob"""
        if self._sha is None:
            self._sha = hashlib.sha256(self.flat_params().tobytes()).hexdigest()
        if version is None:
            version = _lib.load().edmp_unet_blob_layout_version()
        return os.path.join(self.model_name, "edmp_cache", "unet_%s_%s_v%d_%s.blob" % (
            self._sha[:16], self.precision, version, self.plan_class(max_rows)))

    def _drop_stale_blobs(self, keep):
        """blobs of this checkpoint / precision / plan class written by another layout version are dead weight"""
        folder, name = os.path.split(keep)
        head, tail = name.split("_v")[0], name.split("_v")[1].split("_", 1)[1]
        for f in os.listdir(folder):
            if f != name and f.startswith(head + "_v") and f.endswith("_" + tail):
                try:
                    os.remove(os.path.join(folder, f))
                except OSError:
                    pass

    def engine(self, rows=1):
        """Opaque edmp_unet* with a workspace for at least ``rows`` rows."""
        dev = _lib.require_cuda(self.device)
        lib = _lib.load()
        if self._handle is None or rows > self._max_rows:
            self._release()
            self._max_rows = max(self._max_rows, int(rows))
            dims = (ctypes.c_int * len(self.dims))(*self.dims)
            handle = ctypes.c_void_p()
            path = self.blob_path(self._max_rows) if self._cache_enabled() else None
            with torch.cuda.device(dev):
                if path is not None and os.path.exists(path):
                    blob = np.fromfile(path, dtype=np.uint8)
                    rc = lib.edmp_unet_create_from_blob(blob.ctypes.data_as(ctypes.c_void_p), blob.size, self._max_rows,
                                                        ctypes.byref(handle))
                    if rc == 0:
                        self._handle, self.engine_source = handle, "blob"
                        return self._handle
                    warnings.warn("edmp_b200: dropping the cached packed-weight blob %s (%s)" % (
                        path, lib.edmp_last_error().decode()), RuntimeWarning)
                    try:
                        os.remove(path)
                    except OSError:
                        pass
                    handle = ctypes.c_void_p()
                flat = self.flat_params()
                expect = lib.edmp_unet_param_count(dims, len(self.dims))
                if expect != flat.size:
                    raise _lib.EdmpError("state_dict has %d floats, library expects %d" % (flat.size, expect))
                create = lib.edmp_unet_pack if path is not None else lib.edmp_unet_create
                _lib.check(create(flat.ctypes.data_as(ctypes.c_void_p), flat.size, dims, len(self.dims),
                                  _lib.PRECISIONS[self.precision], self._max_rows, ctypes.byref(handle)),
                           "edmp_unet_pack" if path is not None else "edmp_unet_create")
                self._handle, self.engine_source = handle, "state_dict"
                if path is not None:
                    try:
                        n = lib.edmp_unet_blob_bytes(handle)
                        blob = np.empty(n, dtype=np.uint8)
                        _lib.check(lib.edmp_unet_blob_read(handle, blob.ctypes.data_as(ctypes.c_void_p), n),
                                   "edmp_unet_blob_read")
                        os.makedirs(os.path.dirname(path), exist_ok=True)
                        tmp = path + ".tmp%d" % os.getpid()
                        blob.tofile(tmp)
                        os.replace(tmp, path)            # atomic: a concurrent reader sees the old file or the new one
                        self._drop_stale_blobs(path)
                        self.engine_source = "packed"
                    except OSError as e:                 # read-only model directory: run without the cache
                        warnings.warn("edmp_b200: cannot write the packed-weight cache (%s)" % e, RuntimeWarning)
        return self._handle

    def forward(self, x, t):
        """x: [B, 7, 50] float32 tensor, t: [1] tensor (or number) -> eps [B, 7, 50] on the device."""
        dev = _lib.require_cuda(self.device)
        x = torch.as_tensor(x).to(dev, torch.float32).contiguous()
        if x.dim() != 3 or x.shape[1] != 7 or x.shape[2] != 50:
            raise ValueError("expected x of shape [B, 7, 50], got %s" % (tuple(x.shape),))
        tv = float(t.reshape(-1)[0]) if torch.is_tensor(t) else float(t)
        if tv != int(tv) or not 1 <= int(tv) <= 255:
            raise ValueError("t must be an integer step in 1..255 (time embeddings are tabulated)")
        eps = torch.empty_like(x)
        with torch.cuda.device(dev):
            handle = self.engine(x.shape[0])
            _lib.check(_lib.load().edmp_unet_forward(handle, ctypes.c_void_p(x.data_ptr()), int(tv), x.shape[0],
                                                     ctypes.c_void_p(eps.data_ptr()), _lib.stream_ptr()),
                       "edmp_unet_forward")
        return eps

    __call__ = forward

    def check_range(self):
        """Raises if an activation of a forward since the last check left the IEEE-half operand range of the f16x3 /
        f16 modes (+-65504; the results would be NaN).  Synchronises the stream; the other modes never raise."""
        if self._handle is None:
            return
        flag = ctypes.c_int(0)
        with torch.cuda.device(torch.device(self.device)):
            _lib.check(_lib.load().edmp_unet_range_status(self._handle, ctypes.byref(flag), _lib.stream_ptr()),
                       "edmp_unet_range_status")
        if flag.value:
            raise _lib.EdmpError("an activation exceeded the IEEE-half operand range (+-65504) of precision %r: the "
                                 "output is not finite. Use precision='tf32x3' (or 'bf16x3' / 'fp32') for this "
                                 "checkpoint." % self.precision)

    def read_activation(self, name, rows):
        """[rows, C, L] activation of the last forward, named like the reference module path."""
        lib = _lib.load()
        dev = _lib.require_cuda(self.device)
        C, L = ctypes.c_int(), ctypes.c_int()
        with torch.cuda.device(dev):
            _lib.check(lib.edmp_unet_read_activation(self._handle, name.encode(), rows, None, ctypes.byref(C),
                                                     ctypes.byref(L), _lib.stream_ptr()), "read_activation")
            out = torch.empty(rows, C.value, L.value, device=dev, dtype=torch.float32)
            _lib.check(lib.edmp_unet_read_activation(self._handle, name.encode(), rows,
                                                     ctypes.c_void_p(out.data_ptr()), ctypes.byref(C),
                                                     ctypes.byref(L), _lib.stream_ptr()), "read_activation")
        return out
