"""Minimal attribute-access config tree.  The reference builds quantizers / observers / sparsers
from a yacs ``CfgNode`` sub-tree (sparsebit/quantization/quant_config.py:6-48); the plugin classes
here only use attribute access, so a yacs node, this ``Node`` or any namespace works."""


class Node(dict):
    def __init__(self, mapping=None, **kw):
        super().__init__()
        for k, v in dict(mapping or {}, **kw).items():
            self[k] = Node(v) if isinstance(v, dict) and not isinstance(v, Node) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None

    def __setattr__(self, k, v):
        self[k] = v


def quantizer_config(qscheme, bit, target, observer="minmax", layout="NCHW", alpha=0.001, qtype="uniform", disable=False,
                     pact_alpha=10, ema_ratio=0.9, aciq_distribution="GAUS"):
    """The ``config.W`` / ``config.A`` sub-tree + TARGET that ``build_quantizer`` receives
    (quant_model.py:97-137).  ``target``: "weight" | "feature"."""
    from .quantization.common import QuantTarget

    obs = {"TYPE": observer, "PERCENTILE": {"ALPHA": alpha}, "ACIQ": {"DISTRIBUTION": aciq_distribution}}
    quantizer = {"TYPE": qtype, "DISABLE": disable, "BIT": bit}
    if target == "feature":
        obs["LAYOUT"] = layout
        obs["MOVING_AVERAGE"] = {"EMA_RATIO": ema_ratio}   # quant_config.py:41-42
        quantizer["PACT"] = {"ALPHA_VALUE": pact_alpha}    # quant_config.py:35-36
    return Node(
        QSCHEME=qscheme,
        QUANTIZER=quantizer,
        OBSERVER=obs,
        TARGET=(QuantTarget.WEIGHT if target == "weight" else QuantTarget.FEATURE,),
    )


def sparser_config(ratio, stype="unstructed", strategy="l1norm"):
    return Node(SPARSER={"TYPE": stype, "STRATEGY": strategy, "RATIO": ratio})
