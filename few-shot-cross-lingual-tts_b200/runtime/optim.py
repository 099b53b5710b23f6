"""Fused clip + Adam + LR schedule on the flat parameter / gradient buffers (csrc/optim.cu).

Replaces, for the FastSpeech2 path, what the reference gets from pytorch_lightning + torch.optim:
`gradient_clip_val` (main.py:104-110), `torch.optim.Adam` (lightning/optimizer.py:5-16) and the
LambdaLR schedules (lightning/scheduler.py:5-62).  Parameters are re-homed into one flat fp32 buffer
with the same element offsets as `GradBuckets.flat`, so one kernel updates all 34.5 M of them; the step
counter lives on the device, which makes `step()` capturable in the training-step CUDA graph.
"""
import ctypes

import torch

from .. import _cabi

_SCHED = {"none": 0, "sqrt": 1, "const": 2}


class FusedAdam:
    def __init__(self, buckets, train_config=None, lr=None, betas=None, eps=None, weight_decay=None,
                 max_grad_norm=None, scheduler_type=None):
        """`train_config` is the reference's train yaml dict (config/train/baseline.yaml: optimizer.{betas,
        eps, weight_decay, grad_clip_thresh, warm_up_step, anneal_steps, anneal_rate}, scheduler_type);
        explicit keyword arguments override it."""
        opt = dict((train_config or {}).get("optimizer", {}))
        self.lr0 = float(lr if lr is not None else opt.get("lr", 0.001))
        b = betas if betas is not None else opt.get("betas", (0.9, 0.999))
        self.beta1, self.beta2 = float(b[0]), float(b[1])
        self.eps = float(eps if eps is not None else opt.get("eps", 1e-8))
        self.weight_decay = float(weight_decay if weight_decay is not None else opt.get("weight_decay", 0.0))
        self.max_norm = float(max_grad_norm if max_grad_norm is not None else opt.get("grad_clip_thresh", 0.0))
        st = scheduler_type if scheduler_type is not None else (train_config or {}).get(
            "scheduler_type", "sqrt" if train_config else "none")
        self.sched_type = _SCHED[st]
        self.warmup = int(opt.get("warm_up_step", 0))
        self.anneal_steps = [int(s) for s in opt.get("anneal_steps", [])]
        self.anneal_rate = float(opt.get("anneal_rate", 1.0))
        if len(self.anneal_steps) > 8:
            raise ValueError("at most 8 anneal steps")
        self._anneal_c = (ctypes.c_int32 * 8)(*(self.anneal_steps + [0] * (8 - len(self.anneal_steps))))

        self.buckets = buckets
        g = buckets.flat
        if not g.is_cuda:
            raise RuntimeError("FusedAdam runs on sm_100a CUDA kernels only (there is no CPU path)")
        # re-home the parameters: same offsets as their gradients in buckets.flat
        self.flat_param = buckets.rehome_parameters()
        self.exp_avg = torch.zeros_like(g)
        self.exp_avg_sq = torch.zeros_like(g)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=g.device)
        self.gnorm_sq = torch.zeros(1, dtype=torch.float32, device=g.device)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=g.device)  # last un-clipped global norm

    def step(self):
        """Clip by global norm, Adam update, advance the step counter.  Stream-ordered, no host sync."""
        L = _cabi.lib()
        st = torch.cuda.current_stream().cuda_stream
        g = self.buckets.flat
        n = g.numel()
        if self.max_norm > 0:
            _cabi.check(L.fs2_sumsq_f32(g.data_ptr(), n, self.gnorm_sq.data_ptr(), st), "sumsq")
        _cabi.check(L.fs2_adam_step_f32(
            self.flat_param.data_ptr(), g.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), n,
            self.gnorm_sq.data_ptr(), self.step_dev.data_ptr(), self.lr0, self.beta1, self.beta2, self.eps,
            self.weight_decay, self.max_norm, self.sched_type, self.warmup,
            ctypes.cast(self._anneal_c, ctypes.c_void_p), len(self.anneal_steps), self.anneal_rate, st), "adam_step")
        _cabi.check(L.fs2_optim_advance(self.step_dev.data_ptr(), self.gnorm_sq.data_ptr(),
                                        self.grad_norm.data_ptr(), st), "optim_advance")

    # -- checkpointing: torch.optim.Adam's own state_dict format ------------------------------------------------
    # (the reference's Lightning checkpoints store `optimizer_states` = [torch.optim.Adam.state_dict()] for
    # Adam(model.parameters()), lightning/optimizer.py:5-16, main.py resume_from_checkpoint): per-parameter entries
    # keyed by the parameter's index in the list the optimizer was built over, tensors in the parameter's LOGICAL
    # shape -- independent of the flat bucket order / the [Co][k][Ci] storage of Conv1d weights used internally.
    def _logical_view(self, buf, p):
        off = (p.main_grad.data_ptr() - self.buckets.flat.data_ptr()) // 4
        flat = buf[off: off + p.numel()]
        if p.main_grad.is_contiguous():
            return flat.view(p.shape)
        Co, Ci, k = p.shape
        return flat.view(Co, k, Ci).permute(0, 2, 1)

    def state_dict(self):
        index = {id(p): i for i, p in enumerate(self.buckets.all_params)}
        step = self.step_dev.to(torch.float32).reshape(()).clone()
        state = {}
        if int(self.step_dev) > 0:  # torch.optim.Adam has no state before its first step
            for p in self.buckets.params:
                state[index[id(p)]] = {"step": step.clone(),
                                       "exp_avg": self._logical_view(self.exp_avg, p).contiguous().clone(),
                                       "exp_avg_sq": self._logical_view(self.exp_avg_sq, p).contiguous().clone()}
        group = {"lr": self.lr0, "betas": (self.beta1, self.beta2), "eps": self.eps,
                 "weight_decay": self.weight_decay, "amsgrad": False, "maximize": False,
                 "params": list(range(len(self.buckets.all_params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        """Accepts the format above, i.e. also the `optimizer_states[0]` of a reference checkpoint."""
        if "state" not in sd or "param_groups" not in sd:
            raise ValueError("FusedAdam.load_state_dict: expected a torch.optim.Adam-style state_dict "
                             "({'state': {index: {step, exp_avg, exp_avg_sq}}, 'param_groups': [...]})")
        n_all = len(self.buckets.all_params)
        listed = sd["param_groups"][0]["params"]
        if len(listed) != n_all:
            raise ValueError("FusedAdam.load_state_dict: the checkpoint optimises %d parameters, this model has %d"
                             % (len(listed), n_all))
        index = {id(p): i for i, p in enumerate(self.buckets.all_params)}
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        steps = set()
        for p in self.buckets.params:
            ent = sd["state"].get(listed[index[id(p)]])
            if ent is None:
                continue
            for key, buf in (("exp_avg", self.exp_avg), ("exp_avg_sq", self.exp_avg_sq)):
                t = ent[key]
                if tuple(t.shape) != tuple(p.shape):
                    raise ValueError("FusedAdam.load_state_dict: %s of parameter %d has shape %s, expected %s"
                                     % (key, index[id(p)], tuple(t.shape), tuple(p.shape)))
                self._logical_view(buf, p).copy_(t)
            steps.add(int(ent["step"]))
        if len(steps) > 1:
            raise ValueError("FusedAdam.load_state_dict: parameters with different step counts %s" % sorted(steps))
        self.step_dev.fill_(steps.pop() if steps else 0)
