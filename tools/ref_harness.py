"""Import the UNMODIFIED DSP-SLAM reference (read-only at /root/reference) on a CPU-only box.

Test/fixture infrastructure only -- never imported by the product path.  The reference
hard-codes `.cuda()` and needs three optional third-party modules that are not installed
here; we inject inert stand-ins so that
`reconstruct.optimizer.Optimizer.reconstruct_object` runs on PyTorch-CPU exactly as written.

Nothing from the reference is copied: this file only arranges for it to be importable.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("DSP_SLAM_REF", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "reconstruct", "optimizer.py"))


class _AttrDict(dict):
    """Minimal stand-in for addict.Dict: attribute access, nested dict wrapping, __missing__ hook."""

    def __init__(self, *a, **kw):
        super().__init__()
        for k, v in dict(*a, **kw).items():
            self[k] = self._wrap(v)

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, _AttrDict):
            return cls(v)
        return v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            return self.__missing__(k)

    def __setattr__(self, k, v):
        self[k] = self._wrap(v)

    def __missing__(self, k):
        v = type(self)()
        self[k] = v
        return v


def install_shims():
    """Stub addict/plyfile/skimage.measure and make .cuda() a no-op when no GPU is present."""
    import torch

    if "addict" not in sys.modules:
        m = types.ModuleType("addict")
        m.Dict = _AttrDict
        sys.modules["addict"] = m
    if "plyfile" not in sys.modules:
        m = types.ModuleType("plyfile")
        m.PlyData = object
        m.PlyElement = object
        sys.modules["plyfile"] = m
    if "skimage" not in sys.modules:
        sk = types.ModuleType("skimage")
        me = types.ModuleType("skimage.measure")
        sk.measure = me
        sys.modules["skimage"] = sk
        sys.modules["skimage.measure"] = me
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        torch.cuda.synchronize = lambda *a, **k: None
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


def load():
    """Returns a namespace with the reference modules (optimizer, loss, loss_utils, decoder, utils)."""
    install_shims()
    import importlib

    ns = types.SimpleNamespace()
    ns.loss_utils = importlib.import_module("reconstruct.loss_utils")
    ns.loss = importlib.import_module("reconstruct.loss")
    ns.utils = importlib.import_module("reconstruct.utils")
    ns.optimizer = importlib.import_module("reconstruct.optimizer")
    ns.decoder = importlib.import_module("deep_sdf.deep_sdf_decoder")
    return ns


def load_config(name):
    """configs/config_kitti.json etc. through the reference's own get_configs."""
    ns = load()
    return ns.utils.get_configs(os.path.join(REF_ROOT, "configs", name))
