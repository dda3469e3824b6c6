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
        self._beta_state = {}      # net name -> (exp_avg, exp_avg_sq) of the net's swish beta vector

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
            self._step_betas(name, step, lr, group)          # (before the net's kernel, which increments the step counter)
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

    def _step_betas(self, name, step, lr, group):
        """Adam on the net's swish beta vector (rllib's Swish parameter; at most one scalar per layer): a handful of tensor ops
        on an 8-element tensor, same arithmetic and the same device-side step counter as the fused kernel."""
        betas = getattr(self.model, "_betas", {}).get(name)
        if not betas:
            return
        beta, dbeta = betas
        layers = self.model.net(name).fc_layers()
        acts = [m._model[1] if len(m._model) > 1 else None for m in layers]
        if not any(a is not None and hasattr(a, "_beta") and a._beta.grad is not None for a in acts):
            return
        if name not in self._beta_state:
            self._beta_state[name] = (torch.zeros_like(beta), torch.zeros_like(beta))
        m, v = self._beta_state[name]
        b1, b2 = group["betas"]
        g = dbeta if not group["weight_decay"] else dbeta + group["weight_decay"] * beta
        t = step + 1.0
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1 = 1 - torch.pow(torch.full_like(t, b1), t)
        bc2_sqrt = (1 - torch.pow(torch.full_like(t, b2), t)).sqrt()
        beta.sub_((m / (v.sqrt() / bc2_sqrt + group["eps"])) * (lr / bc1))
        for l, a in enumerate(acts):
            if a is not None and hasattr(a, "_beta") and a._beta not in self.state:
                self.state[a._beta] = {"step": step, "exp_avg": m[l], "exp_avg_sq": v[l]}

    def prepare(self, names):
        """Allocate the state of the given nets now (a captured step must not allocate)."""
        for name in names:
            self._net_state(name)
            betas = getattr(self.model, "_betas", {}).get(name)
            if betas and name not in self._beta_state:
                self._beta_state[name] = (torch.zeros_like(betas[0]), torch.zeros_like(betas[0]))

    # ---- resume support (the reference saves no optimizer state: a resumed run restarts the moments) ----
    def flat_state_dict(self):
        """{net name: {"exp_avg", "exp_avg_sq", "step"}} on the CPU + the learning rate the schedulers drive."""
        return {"nets": {n: {"exp_avg": a.detach().cpu().clone(), "exp_avg_sq": b.detach().cpu().clone(), "step": c.detach().cpu().clone()}
                         for n, (a, b, c) in self._flat_state.items()},
                "betas": {n: {"exp_avg": a.detach().cpu().clone(), "exp_avg_sq": b.detach().cpu().clone()} for n, (a, b) in self._beta_state.items()},
                "lr": [float(g["lr"]) for g in self.param_groups]}

    def load_flat_state_dict(self, sd):
        for n, st in sd["nets"].items():
            a, b, c = self._net_state(n)
            a.copy_(st["exp_avg"]); b.copy_(st["exp_avg_sq"]); c.copy_(st["step"])
        for n, st in sd.get("betas", {}).items():
            betas = getattr(self.model, "_betas", {}).get(n)
            if betas:
                self._beta_state[n] = (st["exp_avg"].to(betas[0].device), st["exp_avg_sq"].to(betas[0].device))
        for g, lr in zip(self.param_groups, sd["lr"]):
            g["lr"] = lr
