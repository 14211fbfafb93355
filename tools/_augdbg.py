import sys, os, importlib
import numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import __graft_entry__ as g
pkg = g.load_package()
import test_gpu_parity as T
src = open("/root/repo/tests/test_gpu_parity.py").read()
a = src.index("def test_augmented_dtsdf_unusual_transforms")
b = src.index("def test_identity_augmentation_is_the_plain_path")
body = src[a:b].replace("assert _same(got[i], cv[i]), \"frame %d layout %d\" % (i, layout)",
   "print('frame', i, 'layout', layout, 'same', _same(got[i], cv[i]), 'ndiff', int((got[i].view(np.uint32) != cv[i].view(np.uint32)).sum()), 'first', np.argwhere(got[i].view(np.uint32) != cv[i].view(np.uint32))[:3].tolist(), got[i][got[i].view(np.uint32) != cv[i].view(np.uint32)][:4], cv[i][got[i].view(np.uint32) != cv[i].view(np.uint32)][:4])")
ns = dict(T.__dict__)
exec(body[:body.rindex("\n\n")] if False else body, ns)
ns["test_augmented_dtsdf_unusual_transforms"]()
