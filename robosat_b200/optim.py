"""Adam with torch.optim.Adam's defaults and state_dict layout (train.py:81, 156-160), stepping every parameter
in ONE kernel over a flat fp32 arena instead of a Python loop over 170 tensors.

Parameters are re-homed into views of a single contiguous buffer (same for grads and both moments), so
`rsb_adam_step` runs once per step. `state_dict()` / `load_state_dict()` speak torch.optim.Adam's format:
{"state": {idx: {"step", "exp_avg", "exp_avg_sq"}}, "param_groups": [...]}; parameters that never received a
gradient (the unused `resnet.fc.*`) get no state entry, exactly like torch.
"""

import torch

from robosat_b200 import _lib


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, guard=True):
        """guard=True (default): a step whose gradients contain inf / NaN (an fp16 activation-gradient overflow under the loss
        scale) is skipped on the device -- weights and moments stay untouched -- and counted; `LossScaler` halves the scale."""
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False)
        super().__init__(params, defaults)
        ps = [p for g in self.param_groups for p in g["params"]]
        assert ps and all(p.is_cuda and p.dtype == torch.float32 for p in ps), "robosat_b200.optim.Adam needs fp32 CUDA parameters"
        dev = ps[0].device
        n = sum(p.numel() for p in ps)
        self._flat_p = torch.empty(n, dtype=torch.float32, device=dev)
        self._flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        self._flat_m = torch.zeros(n, dtype=torch.float32, device=dev)
        self._flat_v = torch.zeros(n, dtype=torch.float32, device=dev)
        self._spans = []
        off = 0
        for p in ps:
            k = p.numel()
            self._flat_p[off:off + k].copy_(p.detach().reshape(-1))
            p.data = self._flat_p[off:off + k].view(p.shape)
            p.grad = self._flat_g[off:off + k].view(p.shape)
            self._spans.append((off, k))
            off += k
        self._params = ps
        self._touched = [False] * len(ps)
        self._step = 0
        self._guard = torch.zeros(4, dtype=torch.int32, device=dev) if guard else None

    @property
    def flat_grad(self):
        """all gradients as ONE contiguous fp32 tensor (what data-parallel training all-reduces in a single call)"""
        return self._flat_g

    def zero_grad(self, set_to_none=False):
        self._flat_g.zero_()
        for p, (off, k) in zip(self._params, self._spans):
            if p.grad is None or p.grad.data_ptr() != self._flat_g[off:off + k].data_ptr():
                p.grad = self._flat_g[off:off + k].view(p.shape)

    @torch.no_grad()
    def step(self, closure=None):
        assert closure is None
        self._step += 1
        g = self.param_groups[0]
        if self._guard is not None:
            _lib.check(_lib.load().rsb_adam_step_guarded(self._flat_p.data_ptr(), self._flat_g.data_ptr(), self._flat_m.data_ptr(),
                                                         self._flat_v.data_ptr(), self._flat_p.numel(), g["lr"], g["betas"][0], g["betas"][1], g["eps"],
                                                         self._step, self._guard.data_ptr(), _lib.current_stream_ptr()), "rsb_adam_step_guarded")
            return
        _lib.check(_lib.load().rsb_adam_step(self._flat_p.data_ptr(), self._flat_g.data_ptr(), self._flat_m.data_ptr(), self._flat_v.data_ptr(),
                                             self._flat_p.numel(), g["lr"], g["betas"][0], g["betas"][1], g["eps"], self._step,
                                             _lib.current_stream_ptr()), "rsb_adam_step")

    @property
    def guard_state(self):
        """device int32[4]: [scratch flag, steps skipped, overflow flag of the last finished step, steps seen] (None without guard)"""
        return self._guard

    def skipped_steps(self):
        """number of steps skipped because of non-finite gradients (synchronises; for logging / tests)"""
        return 0 if self._guard is None else int(self._guard[1].item())

    def mark_used(self, mask):
        """mask[i] = parameter i takes part in the graph (gets gradients); others keep no optimiser state, like torch."""
        self._touched = list(mask)

    def state_dict(self):
        state = {}
        for i, (used, (off, k), p) in enumerate(zip(self._touched, self._spans, self._params)):
            if used and self._step > 0:
                state[i] = {"step": torch.tensor(float(self._step)), "exp_avg": self._flat_m[off:off + k].view(p.shape).clone(),
                            "exp_avg_sq": self._flat_v[off:off + k].view(p.shape).clone()}
        groups = [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]
        groups[0]["params"] = list(range(len(self._params)))
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        for i, st in sd["state"].items():
            off, k = self._spans[int(i)]
            self._flat_m[off:off + k].copy_(st["exp_avg"].reshape(-1))
            self._flat_v[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
            self._touched[int(i)] = True
            self._step = int(st["step"]) if not torch.is_tensor(st["step"]) else int(st["step"].item())
        for g, saved in zip(self.param_groups, sd["param_groups"]):
            for key in ("lr", "betas", "eps"):
                if key in saved:
                    g[key] = tuple(saved[key]) if key == "betas" else saved[key]


class LossScaler:
    """Dynamic loss scale for the fp16 activation gradients of `robosat_b200.UNet` in training mode (the reference trains in fp32
    and needs none; train.py:179-188). The overflow decision itself is taken on the device by the guarded Adam step; this class
    only ADAPTS the scale, reading the device flag through an asynchronous copy one step late -- no synchronisation on the
    training stream. Overflow -> scale / 2 (never below `min_scale`); `growth_interval` clean steps -> scale * 2 (capped)."""

    def __init__(self, net, optimizer, init_scale=4096.0, growth_interval=1000, min_scale=1.0, max_scale=65536.0):
        self.net = net.module if hasattr(net, "module") else net
        self.opt = optimizer
        self.scale, self.growth_interval, self.min_scale, self.max_scale = float(init_scale), growth_interval, min_scale, max_scale
        self._clean = 0
        self._host = torch.zeros(4, dtype=torch.int32, pin_memory=True)
        self._event = None
        self._seen = 0
        self.overflows = 0
        self._apply()

    def _apply(self):
        self.net.loss_scale = self.scale
        for eng in getattr(self.net, "_train_engines", {}).values():
            eng.loss_scale = self.scale

    def update(self):
        """call once per training step, after optimizer.step()"""
        if self.opt.guard_state is None:
            return
        if self._event is not None and self._event.query():
            steps, flag = int(self._host[3]), int(self._host[2])
            if steps > self._seen:  # a finished step we have not accounted for yet
                self._seen = steps
                if flag:
                    self.overflows += 1
                    self._clean = 0
                    self.scale = max(self.min_scale, self.scale * 0.5)
                    self._apply()
                else:
                    self._clean += 1
                    if self._clean >= self.growth_interval and self.scale < self.max_scale:
                        self._clean = 0
                        self.scale = min(self.max_scale, self.scale * 2.0)
                        self._apply()
            self._event = None
        if self._event is None:
            self._host.copy_(self.opt.guard_state, non_blocking=True)
            self._event = torch.cuda.Event()
            self._event.record()
