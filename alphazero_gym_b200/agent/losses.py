"""Loss objects with the reference's constructor keywords (alphazero/agent/losses.py:154-209 A0CLoss, :329-429 A0CLossTuned), the
`_target_`s of config/loss/*.yaml.  They only carry the hyper-parameters; the loss arithmetic is train.A0CLoss (device tensors,
CUDA-graph capturable), pinned against the unmodified reference by tests/test_train_step.py."""
from __future__ import annotations

from ..train import LossConfig


class _Loss:
    cfg: LossConfig

    def to(self, device):  # the reference moves the loss module to the device (agents.py:83)
        self.device = device
        return self


class A0CLoss(_Loss):
    def __init__(self, tau: float, policy_coeff: float, alpha: float, value_coeff: float, reduction: str):
        self.cfg = LossConfig(tuned=False, tau=tau, policy_coeff=policy_coeff, value_coeff=value_coeff, alpha=alpha, reduction=reduction)


class A0CLossTuned(_Loss):
    def __init__(self, action_dim: int, alpha_init: float, lr: float, tau: float, policy_coeff: float, value_coeff: float, reduction: str,
                 grad_clip: float, device: str):
        self.cfg = LossConfig(tuned=True, tau=tau, policy_coeff=policy_coeff, value_coeff=value_coeff, alpha=alpha_init, reduction=reduction,
                              alpha_lr=lr, action_dim=action_dim, alpha_grad_clip=grad_clip)


class AlphaZeroLoss(_Loss):
    def __init__(self, *a, **kw):
        raise NotImplementedError("DiscreteAgent + AlphaZeroLoss does not run upstream either (agents.py:378-381 hands a Categorical to "
                                  "F.cross_entropy); use A0CLoss, the loss config/run_discrete.yaml ships")
