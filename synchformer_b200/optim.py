"""Optimiser side of the training step on the sm_100a kernels ("fused CE, grad-clip + Adam on 22.6 M params").

`cross_entropy(logits, targets)`: `F.cross_entropy` (mean reduction) with its gradient produced in the same launch
(`Synchformer.compute_loss`, sync_model.py:91-99).

`FusedAdam`: `torch.optim.Adam` semantics (L2 weight decay, no amsgrad) over all parameters in ONE kernel launch, optionally fused with what
`make_backward_and_optim_step` (scripts/train_utils.py:373-386) does around it - the GradScaler's unscale, `clip_grad_norm_(max_norm)` and
the skip-step-on-inf/nan rule - all driven by a device-side gradient norm, so the host never synchronises:

    opt = FusedAdam(model.parameters(), lr, betas, eps, weight_decay)
    # drop-in use (the harness keeps its GradScaler / clip_grad_norm_ calls):   scaler.step(opt)  or  opt.step()
    # fused use (replaces unscale_ + clip_grad_norm_ + step):                   norm, found_inf = opt.step(grad_scale=s, max_norm=1.0)

Parameters and gradients must be contiguous fp32 CUDA tensors.  State (`exp_avg`, `exp_avg_sq`, `step` per parameter as in
torch.optim.Adam - the step counter is one device tensor shared by a group) is created lazily and round-trips through `state_dict()`; every
step bumps the parameters' version counters so the bf16 weight caches of model.py are rebuilt.  No CPU path.
"""
import ctypes
from typing import Optional

import torch

from . import _lib, ops
from ._lib import check


class _CrossEntropyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits: torch.Tensor, targets: torch.Tensor):
        B, C = logits.shape
        loss = torch.empty((1,), device=logits.device, dtype=torch.float32)
        dlogits = torch.empty_like(logits)
        row_loss = torch.empty((B,), device=logits.device, dtype=torch.float32)
        check(_lib.load().sfb_cross_entropy(ops._p(logits), ops._p(targets), B, C, ops._p(loss), ops._p(dlogits), ops._p(row_loss), ops._stream(logits)),
              'sfb_cross_entropy')
        ops._count(2)
        ctx.save_for_backward(dlogits)
        return loss.view(())

    @staticmethod
    def backward(ctx, g: torch.Tensor):
        (dlogits,) = ctx.saved_tensors
        return dlogits * g, None            # g is the scalar upstream gradient (the GradScaler's scale): (B, n_cls) elements


def cross_entropy(logits: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    """mean_b(-log softmax(logits)[b, targets[b]]); logits (B, C) fp32, targets (B,) int64 class indices."""
    ops.require_cuda(logits, 'logits')
    assert logits.dim() == 2 and targets.shape == (logits.shape[0],) and targets.dtype == torch.int64
    return _CrossEntropyFn.apply(logits.float().contiguous(), targets.contiguous())


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError('invalid Adam hyper-parameters')
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self._dev_state = {}

    def _group_state(self, gi: int, device, ps):
        """Device-side scalars of one parameter group.  The completed-step counter lives in `self.state[p]['step']` (ONE 0-dim fp32 device
        tensor shared by all parameters of the group), so it round-trips through `state_dict()` / `load_state_dict()` exactly like
        torch.optim.Adam's per-parameter `step` - the reference logger checkpoints and restores the optimiser (utils/logger.py:146-151)."""
        st = self._dev_state.get(gi)
        if st is None or st['sqnorm'].device != device:
            st = dict(step=None, sqnorm=torch.zeros((1,), device=device, dtype=torch.float32),
                      found_inf=torch.zeros((1,), device=device, dtype=torch.float32))
            self._dev_state[gi] = st
        shared = st['step']
        if shared is None or any(self.state[p].get('step') is not shared for p in ps):
            # first step, new parameters, or a state restored by load_state_dict (possibly torch.optim.Adam's: one tensor per parameter,
            # possibly on the CPU): adopt the largest restored count (they are equal in any state this class or Adam wrote)
            seen = [float(self.state[p]['step']) for p in ps if torch.is_tensor(self.state[p].get('step')) or isinstance(self.state[p].get('step'), (int, float))]
            if shared is not None:
                seen.append(float(shared))
            shared = torch.full((), max(seen) if seen else 0.0, device=device, dtype=torch.float32)
            for p in ps:
                self.state[p]['step'] = shared
            st['step'] = shared
        return st

    def state_dict(self):
        """torch.optim.Optimizer.state_dict with one `step` tensor PER parameter (copies of the group's shared counter), which is the
        layout torch.optim.Adam writes and expects - it increments every parameter's `step` separately."""
        sd = super().state_dict()
        sd['state'] = {k: ({**v, 'step': v['step'].detach().clone()} if torch.is_tensor(v.get('step')) else dict(v)) for k, v in sd['state'].items()}
        return sd

    @staticmethod
    def _bump_versions(ps):
        """The kernel writes through raw pointers, which torch's version counters do not see; everything keyed on `p._version` (the bf16
        GEMM-weight caches of model.py, autograd's saved-tensor checks) must observe the update."""
        setter = getattr(torch._C._autograd, '_unsafe_set_version_counter', None)
        if setter is not None:
            try:
                setter(tuple(ps), tuple(p._version + 1 for p in ps))
                return
            except TypeError:
                pass
        torch._foreach_add_(list(ps), 0.0)            # no-op in value, bumps every version counter

    @torch.no_grad()
    def step(self, closure=None, grad_scale: Optional[float] = None, max_norm: Optional[float] = None):
        """One Adam step.  grad_scale: the loss scale the gradients still carry (None = already unscaled); max_norm: clip the global
        gradient norm of each parameter group to it (None = no clipping).  Returns (grad_norm, found_inf) device tensors of the last
        group when either is given (no sync), else None."""
        if closure is not None:
            raise NotImplementedError('closure is not supported')
        lib = _lib.load()
        chunk = lib.sfb_optim_chunk_elems()
        ret = None
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group['params'] if p.grad is not None]
            if not ps:
                continue
            dev = ps[0].device
            ops.require_cuda(ps[0], 'parameters')
            rows, chunk_tensor, chunk_start = [], [], []
            for k, p in enumerate(ps):
                g = p.grad
                if p.dtype != torch.float32 or g.dtype != torch.float32 or not p.is_contiguous():
                    raise TypeError('FusedAdam needs contiguous fp32 parameters and gradients')
                if not g.is_contiguous():
                    p.grad = g = g.contiguous()
                st = self.state[p]
                if 'exp_avg' not in st:
                    st['exp_avg'], st['exp_avg_sq'] = torch.zeros_like(p), torch.zeros_like(p)
                for k_ in ('exp_avg', 'exp_avg_sq'):          # restored states may sit on another device / dtype
                    if st[k_].device != p.device or st[k_].dtype != torch.float32 or not st[k_].is_contiguous():
                        st[k_] = st[k_].to(device=p.device, dtype=torch.float32).contiguous()
                n = p.numel()
                rows.append([p.data_ptr(), g.data_ptr(), st['exp_avg'].data_ptr(), st['exp_avg_sq'].data_ptr(), n])
                for s in range(0, n, chunk):
                    chunk_tensor.append(k)
                    chunk_start.append(s)
            table = torch.tensor(rows, dtype=torch.int64).to(dev)
            ct = torch.tensor(chunk_tensor, dtype=torch.int32).to(dev)
            cs = torch.tensor(chunk_start, dtype=torch.int64).to(dev)
            n_chunks = len(chunk_tensor)
            gs = self._group_state(gi, dev, ps)
            stream = ops._stream(ps[0])
            fused = grad_scale is not None or max_norm is not None
            if fused:
                partial = torch.empty((n_chunks,), device=dev, dtype=torch.float32)
                check(lib.sfb_grad_sqnorm(ops._p(table), ops._p(ct), ops._p(cs), n_chunks, ops._p(partial), ops._p(gs['sqnorm']), stream), 'sfb_grad_sqnorm')
                ops._count(2)
            else:
                gs['sqnorm'].zero_()
            inv_scale = 1.0 / float(grad_scale) if grad_scale is not None else 1.0
            b1, b2 = group['betas']
            check(lib.sfb_adam_step(ops._p(table), ops._p(ct), ops._p(cs), n_chunks, ops._p(gs['sqnorm']), ops._p(gs['found_inf']), ops._p(gs['step']),
                                    float(group['lr']), float(b1), float(b2), float(group['eps']), float(group['weight_decay']), inv_scale,
                                    float(max_norm) if max_norm is not None else 0.0, stream), 'sfb_adam_step')
            ops._count(2)
            self._bump_versions(ps)
            if fused:
                ret = (gs['sqnorm'].sqrt() * inv_scale, gs['found_inf'])
        return ret
