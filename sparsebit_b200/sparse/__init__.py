"""Host-side mirror of ``sparsebit.sparse`` for the hot path: Sparser registry, the L1-norm
sparser's mask generation and the mask-apply executed every forward."""
from .modules import SBatchNorm2d, SConv2d, SLinear, apply_mask  # noqa: F401
from .sparsers import SPARSERS_MAP, build_sparser, register_sparser  # noqa: F401
