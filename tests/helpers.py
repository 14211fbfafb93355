"""Shared test helpers: package loader and an independent (pure NumPy) batch packer."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
synth = pkg.synth
PRE_DIR = os.path.join(graft.PKG_DIR, "pre")


def pack_simple(frames):
    """Packed batch built without the library (16-byte aligned crop starts, zero padding)."""
    headers = np.stack([np.asarray(f["header"]) for f in frames]).astype(np.int32)
    offs, total = [], 0
    for f in frames:
        d = f["depth"] if "depth" in f else f["data"]
        offs.append(total)
        total += (d.size + 3) & ~3
    packed = np.zeros(max(total, 4), dtype=np.float32)
    for o, f in zip(offs, frames):
        d = f["depth"] if "depth" in f else f["data"]
        packed[o:o + d.size] = d
    return headers, np.array(offs, dtype=np.int64), packed


def import_pre(name):
    """Imports one of the drop-in modules the way the reference's scripts do: by bare name
    with the ``pre/`` directory on sys.path."""
    import importlib
    if PRE_DIR not in sys.path:
        sys.path.insert(0, PRE_DIR)
    return importlib.import_module(name)
