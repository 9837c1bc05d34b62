"""AdamUniform -- same surface as the reference's largesteps/optimize.py, fused on the B200.

Variant of Adam with uniform scaling by the second moment: instead of dividing each component by the square root of
its second moment, all of them are divided by the max (optimize.py:3-41).  The reference spends ~8 eager kernels and a
max-reduction per parameter per step; here a step is two streaming kernels (csrc/ls_adam.cu).
"""
import torch

from . import _native as N


class AdamUniform(torch.optim.Optimizer):
    def __init__(self, params, lr=0.1, betas=(0.9, 0.999)):
        defaults = dict(lr=lr, betas=betas)
        super(AdamUniform, self).__init__(params, defaults)

    def __setstate__(self, state):
        super(AdamUniform, self).__setstate__(state)

    @torch.no_grad()
    def step(self):
        lib = N.lib()
        for group in self.param_groups:
            lr = group['lr']
            b1, b2 = group['betas']
            for p in group["params"]:
                if p.grad is None:
                    raise RuntimeError("AdamUniform.step(): parameter without gradient (the reference dereferences p.grad too)")
                N.require_cuda(p, "parameter")
                if p.dtype != torch.float32 or not p.data.is_contiguous():
                    raise TypeError("AdamUniform (B200) needs contiguous float32 parameters")
                state = self.state[p]
                if len(state) == 0:           # lazy initialization (optimize.py:24-28)
                    state["step"] = 0
                    state["g1"] = torch.zeros_like(p.data)
                    state["g2"] = torch.zeros_like(p.data)
                    state["scratch"] = torch.zeros(4, dtype=torch.int32, device=p.device)
                state["step"] += 1
                t = state["step"]
                grad = p.grad.data
                if grad.dtype != torch.float32 or grad.device != p.device:
                    raise TypeError("gradient must be float32 on the parameter's device")
                grad = grad.contiguous()
                with torch.cuda.device(p.device):
                    N.check(lib.ls_adam_uniform_step(
                        N.ptr(p.data), N.ptr(grad), N.ptr(state["g1"]), N.ptr(state["g2"]), p.numel(),
                        float(lr), float(b1), float(b2), float(1 - b1), float(1 - b2),
                        float(1 - (b1 ** t)), float(1 - (b2 ** t)),
                        N.ptr(state["scratch"]), N.stream_ptr(p.device)), "ls_adam_uniform_step")
