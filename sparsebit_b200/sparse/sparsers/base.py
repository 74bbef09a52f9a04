"""Sparser base (sparsebit/sparse/sparsers/base.py:6-25)."""
from torch import nn


class Sparser(nn.Module):
    STRATEGY = "base"

    def __init__(self, config, opr):
        super().__init__()
        self.config = config
        self.opr = opr
        self.type = config.SPARSER.TYPE
        self.strategy = config.SPARSER.STRATEGY
        self.ratio = config.SPARSER.RATIO

    def calc_mask(self, x):
        raise NotImplementedError

    def set_ratio(self, ratio):
        self.ratio = ratio

    def __repr__(self):
        return "{}, {}, {}".format(self.type, self.strategy, self.ratio)
