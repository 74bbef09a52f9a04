"""``TYPE = "adaround"`` (sparsebit/quantization/quantizers/adaround.py:15-134): weight-only quantizer
that learns, per element, whether to round down or up.  The forward / its gradient / the variable
initialisation are single kernels of the native library (``sb200_adaround_{fwd,bwd,init}``) instead of
the reference's ATen op chain; ``reconstruct_qlayer`` keeps the reference's optimisation recipe."""
import torch
import torch.nn as nn

from ... import ops
from ..common import QuantTarget
from . import Quantizer as BaseQuantizer
from . import register_quantizer

ZETA, GAMMA = 1.1, -0.1  # stretch parameters of the rectified sigmoid (adaround.py:23)


class _SoftRound(torch.autograd.Function):
    """x_dq with h(v) soft rounding; only ``v`` is differentiable (floor() has no gradient and
    scale / zero_point are buffers of this quantizer)."""

    @staticmethod
    def forward(ctx, x, v, scale, zero_point, qdesc):
        x, v = x.contiguous(), v.contiguous()
        scale, zero_point = scale.reshape(-1).contiguous(), zero_point.reshape(-1).contiguous()
        ctx.save_for_backward(x, v, scale, zero_point)
        ctx.qdesc = qdesc
        return ops.adaround_forward(x, v, scale, zero_point, qdesc.qmin, qdesc.qmax, _ch_axis(qdesc), soft=True)

    @staticmethod
    def backward(ctx, gy):
        x, v, scale, zero_point = ctx.saved_tensors
        q = ctx.qdesc
        gv = ops.adaround_backward(x, v, scale, zero_point, gy.contiguous(), q.qmin, q.qmax, _ch_axis(q))
        return None, gv, None, None, None


def _ch_axis(qdesc):
    return qdesc.ch_axis if qdesc.is_perchannel else None


@register_quantizer
class Quantizer(BaseQuantizer):
    TYPE = "adaround"

    def __init__(self, config):
        super().__init__(config)
        assert config.TARGET[0] == QuantTarget.WEIGHT, "AdaRound only supports to quant weights"
        self.zeta, self.gamma = ZETA, GAMMA

    def init_variables(self, x):
        """v such that the soft-round value starts at frac(x / scale) (adaround.py:26-32)."""
        v = ops.adaround_init(x.detach().contiguous(), self.scale.reshape(-1).contiguous(), _ch_axis(self.qdesc))
        self.v = nn.Parameter(v)

    def _qparams_preprocess(self, x):
        assert not self.export_onnx, "please raise an issue in our repo if you need this feature"
        return self.scale, self.zero_point

    def _get_soft_round_values(self):
        # regulariser input of reconstruct_qlayer (adaround.py:40-43); differentiable torch ops
        return torch.clamp(torch.sigmoid(self.v) * (self.zeta - self.gamma) + self.gamma, 0, 1)

    def _forward(self, x, scale, zero_point):
        if self.training:
            return _SoftRound.apply(x, self.v, scale, zero_point, self.qdesc)
        return ops.adaround_forward(x.contiguous(), self.v.detach().contiguous(), scale.reshape(-1).contiguous(),
                                    zero_point.reshape(-1).contiguous(), self.qdesc.qmin, self.qdesc.qmax,
                                    _ch_axis(self.qdesc), soft=False)


class LinearTempDecay:
    """beta schedule of the rounding regulariser: constant during warm-up, then linear from
    start_beta to end_beta (adaround.py:113-134)."""

    def __init__(self, max_steps, rel_start_step, start_beta, end_beta):
        self.max_steps = max_steps
        self.start_step = int(rel_start_step * max_steps)
        self.start_beta = start_beta
        self.end_beta = end_beta

    def __call__(self, step):
        if step < self.start_step:
            return self.start_beta
        progress = (step - self.start_step) / (self.max_steps - self.start_step)
        return self.end_beta + (self.start_beta - self.end_beta) * max(0.0, 1 - progress)


def reconstruct_qlayer(layer, inputs, outputs, batch_size=32, max_steps=20000, beta_range=(20, 2), warmup=0.2, p=2.0,
                       round_loss_weight=1e-3, a_quant=False, print_freq=500, generator=None):
    """Learn ``layer.weight_quantizer.v`` so that the quantized layer reproduces ``outputs`` on
    ``inputs`` (adaround.py:57-110): Adam on v, reconstruction loss |.|^p summed over dim 1, plus the
    annealed rounding regulariser after the warm-up."""
    wq = layer.weight_quantizer
    layer.eval()
    layer.set_quant(w_quant=True, a_quant=a_quant)
    wq.init_variables(layer.weight)
    wq.train()
    optimizer = torch.optim.Adam([wq.v])
    decay = LinearTempDecay(max_steps, warmup, beta_range[0], beta_range[1])
    first_reg_step = int(warmup * max_steps)
    device = layer.weight.device
    inputs, outputs = inputs.to(device), outputs.to(device)
    for step in range(max_steps):
        idx = torch.randperm(inputs.size(0), generator=generator)[:batch_size].to(device)
        optimizer.zero_grad()
        rec_loss = (layer(inputs[idx]) - outputs[idx]).abs().pow(p).sum(1).mean()
        if step < first_reg_step:
            beta = round_loss = 0
        else:
            beta = decay(step)
            soft = wq._get_soft_round_values()
            round_loss = (1 - ((soft - 0.5).abs() * 2).pow(beta)).sum()
        loss = rec_loss + round_loss_weight * round_loss
        loss.backward()
        optimizer.step()
        if print_freq and step % print_freq == 0:
            print("Loss: {:.3f} (rec: {:.3f}, round: {:.3f}) beta={:.2f} step={}".format(
                float(loss), float(rec_loss), float(round_loss), beta, step))
    wq.eval()
