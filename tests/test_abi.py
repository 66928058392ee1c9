"""CPU-only checks of the C-ABI library: it loads, exports every symbol the headers declare,
refuses to run without a GPU (no fallback), and its host-side batch builder honours the layout."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from util import ROOT


@pytest.fixture(scope="module")
def capi():
    from cornetto_b200 import capi as m
    from cornetto_b200.build import ensure_built
    ensure_built()
    return m


def declared_symbols():
    names = set()
    for h in ("corn_gpu.h", "corn_bench.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(corn_(?:gpu|hbatch|bench|shard)_\w+)\s*\(", src))
        names |= set(re.findall(r"\b(sdust(?:_core|_buf_init|_buf_destroy)?)\s*\(", src))     # the reference's sdust.h interface
    return names


def test_exports_every_declared_symbol(capi):
    L = capi.load()
    decl = declared_symbols()
    assert len(decl) >= 30
    assert decl == set(capi.SYMBOLS), decl ^ set(capi.SYMBOLS)
    for s in decl:
        assert hasattr(L, s), s


def test_no_cpu_fallback(capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = capi.load()
    assert L.corn_gpu_device_count() < 0
    with pytest.raises(capi.CornError) as e:
        capi.Context()
    assert "no usable CUDA device" in str(e.value)
    assert L.corn_gpu_strerror(-1).decode().startswith("no usable CUDA device")


def test_hbatch_layout(capi):
    L = capi.load()
    recs = [b"ACGT" * 10, b"", b"T" * 31, b"G" * 32, b"C" * 33, b"A"]
    hb = capi.HostBatch(recs)
    v = hb.view
    assert v.n_rec == len(recs) and v.total_bytes % capi.CORN_ALIGN == 0
    off = np.frombuffer((C.c_uint8 * (8 * v.n_rec)).from_address(v.offset), dtype=np.uint64)
    ln = np.frombuffer((C.c_uint8 * (4 * v.n_rec)).from_address(v.length), dtype=np.uint32)
    seq = np.frombuffer((C.c_uint8 * v.total_bytes).from_address(v.seq), dtype=np.uint8)
    covered = np.zeros(v.total_bytes, dtype=bool)
    for i, r in enumerate(recs):
        assert off[i] % capi.CORN_ALIGN == 0 and ln[i] == len(r)
        assert bytes(seq[off[i]:off[i] + ln[i]]) == r
        covered[off[i]:off[i] + ln[i]] = True
        nxt = off[i + 1] if i + 1 < len(recs) else v.total_bytes
        assert nxt >= off[i] + ln[i] + 1                      # at least one pad byte
    assert not seq[~covered].any()                            # padding is zero
    # room / overflow behaviour
    h = C.c_void_p()
    assert L.corn_hbatch_create(64, 4, C.byref(h)) == 0
    assert L.corn_hbatch_room(h) == 63
    assert L.corn_hbatch_add(h, b"A" * 40, 40) == 0
    assert L.corn_hbatch_room(h) == 0
    assert L.corn_hbatch_add(h, b"A", 1) != 0
    L.corn_hbatch_reset(h)
    assert L.corn_hbatch_room(h) == 63
    L.corn_hbatch_destroy(h)
    hb.close()
