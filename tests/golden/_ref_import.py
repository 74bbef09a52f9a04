"""Import shim that lets the UNMODIFIED reference (``/root/reference``) be imported in the
build container, where ``yacs`` and ``onnx`` are not installed.

Only used by ``make_golden.py`` (fixture generation) -- never at test / bench / run time:
``/root/reference`` does not exist on the GPU box.
"""
import importlib.machinery
import sys
import types

REFERENCE_ROOT = "/root/reference"


class _CfgNode(dict):
    """Minimal stand-in for ``yacs.config.CfgNode`` (attribute-access dict tree)."""

    def __init__(self, init=None, **kw):
        super().__init__()
        for k, v in (init or {}).items():
            self[k] = _CfgNode(v) if isinstance(v, dict) and not isinstance(v, _CfgNode) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        import copy

        return copy.deepcopy(self)

    def defrost(self):
        pass

    def freeze(self):
        pass

    def merge_from_file(self, f):
        raise NotImplementedError

    def merge_from_list(self, lst):
        for k, v in zip(lst[0::2], lst[1::2]):
            node = self
            parts = k.split(".")
            for p in parts[:-1]:
                node = node[p]
            node[parts[-1]] = v


def install_shims():
    if "yacs" not in sys.modules:
        yacs = types.ModuleType("yacs")
        cfgmod = types.ModuleType("yacs.config")
        cfgmod.CfgNode = _CfgNode
        yacs.config = cfgmod
        yacs.__spec__ = importlib.machinery.ModuleSpec("yacs", None)
        cfgmod.__spec__ = importlib.machinery.ModuleSpec("yacs.config", None)
        sys.modules["yacs"] = yacs
        sys.modules["yacs.config"] = cfgmod
    if "onnx" not in sys.modules:
        onnx = types.ModuleType("onnx")
        onnx.__spec__ = importlib.machinery.ModuleSpec("onnx", None)
        sys.modules["onnx"] = onnx
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def make_cfg(qscheme, bit, target, observer="minmax", layout="NCHW", alpha=0.001, qtype="uniform"):
    """Build the config sub-tree ``build_quantizer`` / ``build_observer`` consume
    (sparsebit/quantization/quant_config.py:6-48; quant_model.py:97-137)."""
    obs = {"TYPE": observer, "PERCENTILE": {"ALPHA": alpha}}
    if target == "feature":
        obs["LAYOUT"] = layout
    from sparsebit.quantization.common import QuantTarget

    return _CfgNode(
        {
            "QSCHEME": qscheme,
            "QUANTIZER": {"TYPE": qtype, "DISABLE": False, "BIT": bit},
            "OBSERVER": obs,
            "TARGET": (QuantTarget.WEIGHT if target == "weight" else QuantTarget.FEATURE,),
        }
    )
