"""`PvaeAdam`: torch.optim.Adam's interface and arithmetic (what the reference constructs at torch_models.py:119-122) executed
by the engine: one fused kernel per Linear layer updates the fp32 master, both moments and -- in the same pass -- the bf16
shadow operand the tensor-core kernels read, so there is no separate refresh pass after the step.

Same observable behaviour as the reference's optimizer: it is built over ALL parameters of the model; parameters whose
`.grad` is None (frozen sub-nets, the value branch) are skipped; state (step, exp_avg, exp_avg_sq) appears lazily the first
time a parameter is stepped; `param_groups[0]["lr"]` is what LR schedulers drive.
"""
import torch

from .engine import NET_NAMES


class PvaeAdam(torch.optim.Optimizer):
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0.0 or eps < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or weight_decay < 0.0:
            raise ValueError("invalid Adam hyper-parameters")
        self.model = model
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(list(model.parameters()), defaults)
        self._flat_state = {}      # net name -> (exp_avg flat, exp_avg_sq flat, step scalar, bitmask of initialised layers)

    def _net_state(self, name):
        if name not in self._flat_state:
            g = self.model.flat_grads(name)
            self._flat_state[name] = [torch.zeros_like(g), torch.zeros_like(g), torch.zeros((), dtype=torch.float32, device=g.device)]
        return self._flat_state[name]

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        eng = self.model.engine()
        group = self.param_groups[0]
        lr = group["lr"]
        lr = float(lr) if not torch.is_tensor(lr) else float(lr.item())
        for name in NET_NAMES:
            if name == "value_branch":
                continue
            layers = self.model.net(name).fc_layers()
            mask = 0
            for l, m in enumerate(layers):
                if m.linear.weight.grad is not None and m.linear.bias.grad is not None:
                    mask |= 1 << l
            if not mask:
                continue
            exp_avg, exp_avg_sq, step = self._net_state(name)
            eng.adam_step(name, mask, exp_avg, exp_avg_sq, step, lr, group["betas"][0], group["betas"][1], group["eps"],
                          group["weight_decay"])
            # torch-compatible per-parameter state: views of the flat buffers
            off = 0
            for l, m in enumerate(layers):
                for p in (m.linear.weight, m.linear.bias):
                    k = p.numel()
                    if (mask >> l) & 1 and p not in self.state:
                        self.state[p] = {"step": step, "exp_avg": exp_avg[off:off + k].view_as(p),
                                         "exp_avg_sq": exp_avg_sq[off:off + k].view_as(p)}
                    off += k
        return loss

    # ---- resume support (the reference saves no optimizer state: a resumed run restarts the moments) ----
    def flat_state_dict(self):
        """{net name: {"exp_avg", "exp_avg_sq", "step"}} on the CPU + the learning rate the schedulers drive."""
        return {"nets": {n: {"exp_avg": a.detach().cpu().clone(), "exp_avg_sq": b.detach().cpu().clone(), "step": c.detach().cpu().clone()}
                         for n, (a, b, c) in self._flat_state.items()},
                "lr": [float(g["lr"]) for g in self.param_groups]}

    def load_flat_state_dict(self, sd):
        for n, st in sd["nets"].items():
            a, b, c = self._net_state(n)
            a.copy_(st["exp_avg"]); b.copy_(st["exp_avg_sq"]); c.copy_(st["step"])
        for g, lr in zip(self.param_groups, sd["lr"]):
            g["lr"] = lr
