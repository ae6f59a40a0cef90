"""Oracle: optimizer step.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates torch.optim.Adam as the reference configures it -- Adam(param_groups, lr=0.0, eps=1e-15), no weight decay,
no amsgrad, betas (0.9, 0.999) (renderer/latent_gs_renderer.py:475; stepped at main_train_dimo.py:416-417) -- in the
operation order of torch's single-tensor CPU path (torch/optim/adam.py::_single_tensor_adam): fp32 tensors,
bias corrections as Python doubles.  Pinned bit-exact against torch.optim.Adam itself (the reference's optimizer IS
this torch class) by tests/test_optim_cpu.py.
"""
import torch


def adam_step(params, grads, exp_avgs, exp_avg_sqs, step, lrs, beta1=0.9, beta2=0.999, eps=1e-15):
    """In-place update of `params` (list of fp32 tensors) for update number `step` (1-based); lrs: one per tensor."""
    bias_correction1 = 1 - beta1 ** step
    bias_correction2 = 1 - beta2 ** step
    bias_correction2_sqrt = bias_correction2 ** 0.5
    for p, g, m, v, lr in zip(params, grads, exp_avgs, exp_avg_sqs, lrs):
        m.lerp_(g, 1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        step_size = lr / bias_correction1
        denom = (v.sqrt() / bias_correction2_sqrt).add_(eps)
        p.addcdiv_(m, denom, value=-step_size)
