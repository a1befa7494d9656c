"""Ingest a DeepSDF decoder the way DSP-SLAM hands it over and upload it to the GPU library.

The reference passes the `nn.Module` returned by reconstruct.utils.get_decoder
(deep_sdf/workspace.py:202-223) into Optimizer / MeshExtractor.  This module pulls the Linear
layers out of such a module (or out of a state_dict / npz fixture), folds weight-norm once
(W = g * v / ||v||_row, torch.nn.utils.weight_norm(dim=0) as applied at
deep_sdf/deep_sdf_decoder.py:49-56) and creates the device-resident decoder handle.
"""
import ctypes as C
import json
import numpy as np

from . import _lib


class DecoderWeights:
    """Folded fp32 weights of one decoder: W[k] (out,in), b[k] (out,), latent_in layer, L."""

    def __init__(self, W, b, latent_in, latent_size):
        self.W = [np.ascontiguousarray(w, dtype=np.float32) for w in W]
        self.b = [np.ascontiguousarray(x, dtype=np.float32) for x in b]
        latent_in = tuple(latent_in or ())
        if len(latent_in) > 1:
            raise NotImplementedError("more than one latent_in layer is not supported")
        self.latent_in_layer = int(latent_in[0]) if latent_in else -1
        self.latent_size = int(latent_size)
        if self.W[0].shape[1] != self.latent_size + 3:
            raise ValueError("first layer must take latent_size+3 inputs (xyz_in_all/other variants unsupported)")

    # -- constructors ------------------------------------------------------------------------
    @staticmethod
    def _fold(sd, k):
        if f"lin{k}.weight_v" in sd:
            v = np.asarray(sd[f"lin{k}.weight_v"], dtype=np.float32)
            g = np.asarray(sd[f"lin{k}.weight_g"], dtype=np.float32).reshape(-1, 1)
            norm = np.sqrt(np.sum(v * v, axis=1, keepdims=True, dtype=np.float32))
            return (g * (v / norm)).astype(np.float32)
        if f"lin{k}.parametrizations.weight.original1" in sd:      # new-style parametrization
            v = np.asarray(sd[f"lin{k}.parametrizations.weight.original1"], dtype=np.float32)
            g = np.asarray(sd[f"lin{k}.parametrizations.weight.original0"], dtype=np.float32).reshape(-1, 1)
            norm = np.sqrt(np.sum(v * v, axis=1, keepdims=True, dtype=np.float32))
            return (g * (v / norm)).astype(np.float32)
        return np.asarray(sd[f"lin{k}.weight"], dtype=np.float32)

    @classmethod
    def from_state_dict(cls, sd, latent_size, latent_in=(), **unsupported):
        for key in ("xyz_in_all", "use_tanh", "latent_dropout"):
            if unsupported.get(key):
                raise NotImplementedError(f"decoder option {key} is not supported by the CUDA path")
        sd = {k: (v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)) for k, v in sd.items()}
        if any(k.startswith("bn") for k in sd):
            raise NotImplementedError("LayerNorm decoders (weight_norm=False with norm_layers) are not supported")
        n_lin = 0
        while any(k.startswith(f"lin{n_lin}.") for k in sd):
            n_lin += 1
        W = [cls._fold(sd, k) for k in range(n_lin)]
        b = [np.asarray(sd[f"lin{k}.bias"], dtype=np.float32) for k in range(n_lin)]
        return cls(W, b, latent_in, latent_size)

    @classmethod
    def from_module(cls, module):
        """`module` = deep_sdf.deep_sdf_decoder.Decoder (eval).  Attributes used: latent_in,
        xyz_in_all, use_tanh, latent_dropout, weight_norm, norm_layers, state_dict()."""
        sd = module.state_dict()
        L = None
        for key in ("lin0.weight_v", "lin0.weight", "lin0.parametrizations.weight.original1"):
            if key in sd:
                L = int(sd[key].shape[1]) - 3
        if L is None:
            raise ValueError("module has no lin0 layer")
        if (not getattr(module, "weight_norm", True)) and getattr(module, "norm_layers", None):
            raise NotImplementedError("LayerNorm decoders are not supported")
        return cls.from_state_dict(sd, L, latent_in=tuple(getattr(module, "latent_in", ()) or ()),
                                   xyz_in_all=getattr(module, "xyz_in_all", None),
                                   use_tanh=getattr(module, "use_tanh", False),
                                   latent_dropout=getattr(module, "latent_dropout", False))

    @classmethod
    def from_npz(cls, path):
        d = np.load(path)
        spec = json.loads(bytes(d["spec_json"]).decode())
        sd = {k: d[k] for k in d.files if k != "spec_json"}
        return cls.from_state_dict(sd, spec["latent_size"], latent_in=spec.get("latent_in", ()),
                                   xyz_in_all=spec.get("xyz_in_all"), use_tanh=spec.get("use_tanh"),
                                   latent_dropout=spec.get("latent_dropout"))

    @classmethod
    def coerce(cls, obj):
        if isinstance(obj, cls):
            return obj
        if isinstance(obj, str):
            return cls.from_npz(obj)
        if hasattr(obj, "state_dict"):
            return cls.from_module(obj)
        raise TypeError(f"cannot build decoder weights from {type(obj)}")


class DeviceDecoder:
    """Owns a DspgnDecoder handle (weights resident in HBM in every layout the kernels use)."""

    def __init__(self, weights, device=0):
        lib = _lib.load()
        self.weights = weights
        self.device = device
        spec = _lib.DecoderSpec()
        spec.latent_size = weights.latent_size
        spec.num_linear = len(weights.W)
        if spec.num_linear > _lib.MAX_LINEAR:
            raise ValueError("too many layers")
        for k, w in enumerate(weights.W):
            spec.out_dim[k], spec.in_dim[k] = w.shape
        spec.latent_in_layer = weights.latent_in_layer
        FP = C.POINTER(C.c_float)
        Wp = (FP * len(weights.W))(*[w.ctypes.data_as(FP) for w in weights.W])
        bp = (FP * len(weights.b))(*[x.ctypes.data_as(FP) for x in weights.b])
        h = C.c_void_p()
        _lib.check(lib.dspgn_decoder_create(C.byref(spec), Wp, bp, device, C.byref(h)))
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            _lib.load().dspgn_decoder_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
