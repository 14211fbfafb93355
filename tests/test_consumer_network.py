"""tools/consumer_net.py (harness of BASELINE.json configs[4], not product code) against the unmodified
3D_CNN/network.py: same parameter names and shapes, same forward pass for the same weights and dropout seed.
Runs where the reference tree is present (this container); the GPU box has no /root/reference."""
import importlib.util
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/3D_CNN/network.py"


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present")
def test_harness_network_equals_reference_network():
    ours = _load(os.path.join(ROOT, "tools", "consumer_net.py"), "consumer_net")
    ref = _load(REF, "ref_network")
    torch.manual_seed(0)
    a, b = ours.DenseNet(), ref.DenseNet()
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    assert [tuple(v.shape) for v in sa.values()] == [tuple(v.shape) for v in sb.values()]
    a.load_state_dict(sb)
    x = torch.randn(2, 3, 32, 32, 32)
    for mode in (True, False):            # dropout stays on in eval mode too (the reference never passes `training`)
        a.train(mode), b.train(mode)
        torch.manual_seed(7)
        ya = a(x)
        torch.manual_seed(7)
        yb = b(x)
        assert ya.shape == (2, 63)
        assert torch.equal(ya, yb)
        a.load_state_dict(b.state_dict())     # batch-norm running statistics moved in train mode: keep them equal
