"""GPU: the packed-weight blob -- edmp_unet_pack / edmp_unet_create_from_blob through the C ABI
and the versioned on-disk cache TemporalUNet keeps next to the checkpoint (reference side: temporalunet.py:78-92)."""
import ctypes
import glob
import os
import warnings

import numpy as np
import pytest
import torch

from oracle import unet_oracle, weights

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
DIMS = (32, 64, 128, 256, 512, 512)


def _saved_model(tmp_path, sd, precision="f16x3"):
    from edmp_b200 import TemporalUNet
    d = str(tmp_path / "TemporalUNetModel255_N50")
    m = TemporalUNet(d, 7, 32, DEV, dims=DIMS, precision=precision)
    m.load_state_dict(sd)
    m.losses = np.zeros(1)
    m.save()                                     # weights_latest.pt + losses.npy, the reference's checkpoint layout
    return d


def test_blob_cache_round_trip_and_stale_versions(tmp_path):
    from edmp_b200 import TemporalUNet, _lib
    sd = weights.seeded_state_dict(5)
    d = _saved_model(tmp_path, sd)
    x = torch.randn(70, 7, 50, generator=torch.Generator().manual_seed(1)).to(DEV)
    with torch.no_grad():
        ref = unet_oracle.unet_forward(sd, x.cpu(), 33).numpy()

    m1 = TemporalUNet(d, 7, 32, DEV, dims=DIMS, precision="f16x3")      # loads the checkpoint -> cache on
    e1 = m1(x, 33)
    assert m1.engine_source == "packed"
    path = m1.blob_path(m1._max_rows)
    assert os.path.exists(path) and os.path.getsize(path) > 50e6
    assert "_v%d_" % _lib.load().edmp_unet_blob_layout_version() in os.path.basename(path)

    m2 = TemporalUNet(d, 7, 32, DEV, dims=DIMS, precision="f16x3")      # second process: no repacking
    e2 = m2(x, 33)
    assert m2.engine_source == "blob"
    assert torch.equal(e1, e2)
    assert np.abs(e2.cpu().numpy() - ref).max() <= 2e-5

    # another precision / another batch-size class = another file; an in-memory state_dict never touches the cache
    m3 = TemporalUNet(d, 7, 32, DEV, dims=DIMS, precision="bf16x3")
    m3(x, 33)
    assert m3.engine_source == "packed" and m3.blob_path(m3._max_rows) != path
    m4 = TemporalUNet(str(tmp_path / "fresh"), 7, 32, DEV, dims=DIMS, precision="f16x3")
    m4.load_state_dict(sd)
    m4(x, 33)
    assert m4.engine_source == "state_dict" and not os.path.isdir(str(tmp_path / "fresh" / "edmp_cache"))

    # stale layout version: (a) a file of an older version is ignored and removed when the current one is written,
    # (b) a blob whose header carries another version is refused by the library with rc 4
    blob = np.fromfile(path, dtype=np.uint8)
    old = m1.blob_path(m1._max_rows, version=_lib.load().edmp_unet_blob_layout_version() - 1)
    blob.tofile(old)
    os.remove(path)
    m5 = TemporalUNet(d, 7, 32, DEV, dims=DIMS, precision="f16x3")
    m5(x, 33)
    assert m5.engine_source == "packed" and os.path.exists(path) and not os.path.exists(old)
    stale = blob.copy()
    stale[8:12] = np.frombuffer(np.uint32(_lib.load().edmp_unet_blob_layout_version() + 7).tobytes(), dtype=np.uint8)
    h = ctypes.c_void_p()
    rc = _lib.load().edmp_unet_create_from_blob(stale.ctypes.data_as(ctypes.c_void_p), stale.size, 64, ctypes.byref(h))
    assert rc == 4 and b"stale blob" in _lib.load().edmp_last_error()
    # a corrupted cache file is dropped with a warning and the engine is repacked
    stale.tofile(path)
    m6 = TemporalUNet(d, 7, 32, DEV, dims=DIMS, precision="f16x3")
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        e6 = m6(x, 33)
    assert m6.engine_source == "packed" and any("dropping the cached" in str(i.message) for i in w)
    assert torch.equal(e6, e1)
    assert len(glob.glob(os.path.join(d, "edmp_cache", "*.tmp*"))) == 0


def test_blob_is_refused_for_another_plan(tmp_path):
    """A blob packed for the small-batch plan (narrow column tiles) must not be played into the 8190-row plan (CTA
    pairs, 128-channel tiles): same bytes per item, different tile geometry -> the per-item tags differ -> rc 4."""
    from edmp_b200 import _lib
    lib = _lib.load()
    sd = weights.seeded_state_dict(6)
    flat = torch.cat([v.reshape(-1) for v in sd.values()]).contiguous().numpy()
    dims = (ctypes.c_int * 6)(*DIMS)
    h = ctypes.c_void_p()
    _lib.check(lib.edmp_unet_pack(flat.ctypes.data_as(ctypes.c_void_p), flat.size, dims, 6, _lib.PRECISIONS["f16x3"], 64,
                                  ctypes.byref(h)), "edmp_unet_pack")
    n = lib.edmp_unet_blob_bytes(h)
    blob = np.empty(n, dtype=np.uint8)
    _lib.check(lib.edmp_unet_blob_read(h, blob.ctypes.data_as(ctypes.c_void_p), n), "edmp_unet_blob_read")
    assert lib.edmp_unet_blob_bytes(h) == 0                       # the read released the engine's host copy
    lib.edmp_unet_destroy(h)
    h2 = ctypes.c_void_p()
    assert lib.edmp_unet_create_from_blob(blob.ctypes.data_as(ctypes.c_void_p), n, 64, ctypes.byref(h2)) == 0
    lib.edmp_unet_destroy(h2)
    h3 = ctypes.c_void_p()
    rc = lib.edmp_unet_create_from_blob(blob.ctypes.data_as(ctypes.c_void_p), n, 8190, ctypes.byref(h3))
    assert rc == 4 and b"does not match this engine's plan" in lib.edmp_last_error()
    rc = lib.edmp_unet_create_from_blob(blob[:1000].ctypes.data_as(ctypes.c_void_p), 1000, 64, ctypes.byref(h3))
    assert rc == 4
