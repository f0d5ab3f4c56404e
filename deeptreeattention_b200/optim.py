"""Fused Adam: the optimizer the reference configures around the path, as ONE kernel launch per step.

``FusedAdam(model.parameters(), lr=config["lr"])`` stands in for ``torch.optim.Adam(self.model.parameters(),
lr=self.config["lr"])`` of ``TreeModel.configure_optimizers`` (/root/reference/src/main.py:135-136) and of
``MultiStage.configure_optimizers`` (/root/reference/src/models/multi_stage.py:258-262); it is a
``torch.optim.Optimizer``, so the reference's ``ReduceLROnPlateau`` scheduler (main.py:138-147) drives it unchanged.
Same arithmetic as torch's Adam (betas, eps, L2 weight decay, per-parameter step counts, parameters without a
gradient skipped); the moments of all float32 parameters live in two flat buffers and ``state[p]`` exposes views of
them under torch's own keys (``step``, ``exp_avg``, ``exp_avg_sq``), so ``state_dict()`` checkpoints look like
torch's.  The one float64 parameter of the path (``Hang2020.alpha``) is updated by the same launch.

``capturable=True`` keeps the step counter and the learning rate on the device so that a step captured in a CUDA
graph (``graph.GraphedTrainStep``) replays with correct bias corrections; call ``refresh_lr()`` after a scheduler
changed the learning rate.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List

import torch

from . import _capi

_MAX_TENSORS = 96


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, capturable=False):
        if lr < 0.0 or eps < 0.0 or not (0.0 <= betas[0] < 1.0) or not (0.0 <= betas[1] < 1.0) or weight_decay < 0.0:
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.capturable = bool(capturable)
        self._flat: Dict[int, dict] = {}      # per param group: flat moment buffers, offsets, device scalars

    # ---- state layout -------------------------------------------------------------------------------------------
    def _group_state(self, gi: int, group) -> dict:
        st = self._flat.get(gi)
        if st is not None:
            return st
        f32 = [p for p in group["params"] if p.dtype == torch.float32]
        f64 = [p for p in group["params"] if p.dtype == torch.float64]
        other = [p for p in group["params"] if p.dtype not in (torch.float32, torch.float64)]
        if other or len(f64) > 1 or any(p.numel() != 1 for p in f64):
            raise TypeError("FusedAdam handles float32 parameters plus at most one float64 scalar (Hang2020.alpha)")
        if not f32 and not f64:
            raise ValueError("empty parameter group")
        dev = (f32 or f64)[0].device
        if dev.type != "cuda" or any(p.device != dev for p in group["params"]):
            raise RuntimeError("deeptreeattention_b200 has no CPU path: FusedAdam needs every parameter on one CUDA (sm_100) device")
        offsets, total = {}, 0
        for p in f32:
            offsets[p] = total
            total += p.numel()
        with torch.cuda.device(dev):
            st = {
                "device": dev, "f32": f32, "f64": f64[0] if f64 else None, "offsets": offsets,
                "exp_avg": torch.zeros(max(total, 1), dtype=torch.float32, device=dev),
                "exp_avg_sq": torch.zeros(max(total, 1), dtype=torch.float32, device=dev),
                "moments64": torch.zeros(2, dtype=torch.float64, device=dev),
                "step_dev": torch.zeros((), dtype=torch.int64, device=dev) if self.capturable else None,
                "lr_dev": torch.full((), float(group["lr"]), dtype=torch.float32, device=dev) if self.capturable else None,
            }
        self._flat[gi] = st
        return st

    def _param_state(self, st: dict, p) -> dict:
        s = self.state[p]
        if "step" not in s:
            s["step"] = 0
            if p is st["f64"]:
                s["exp_avg"], s["exp_avg_sq"] = st["moments64"][0], st["moments64"][1]
            else:
                off = st["offsets"][p]
                s["exp_avg"] = st["exp_avg"][off:off + p.numel()].view(p.shape)
                s["exp_avg_sq"] = st["exp_avg_sq"][off:off + p.numel()].view(p.shape)
        return s

    def refresh_lr(self):
        """Copy each group's ``lr`` to its device scalar (capturable mode; call after a scheduler step, outside capture)."""
        for gi, group in enumerate(self.param_groups):
            st = self._flat.get(gi)
            if st is not None and st["lr_dev"] is not None:
                st["lr_dev"].fill_(float(group["lr"]))

    def state_dict(self):
        """torch-Adam-shaped checkpoint.  In capturable mode the step that counts is the DEVICE counter (it advances on every
        CUDA-graph replay, the host-side ``state[p]["step"]`` only when Python runs ``step()``): it is read back here so a
        resumed run continues the bias correction where the saved one stopped."""
        for gi, group in enumerate(self.param_groups):
            st = self._flat.get(gi)
            if st is None or st["step_dev"] is None:
                continue
            step = int(st["step_dev"].item())
            for p in group["params"]:
                if p in self.state and "step" in self.state[p]:
                    self.state[p]["step"] = step
        return super().state_dict()

    def load_state_dict(self, state_dict):
        """Loads a torch-Adam-shaped checkpoint: the moments are copied INTO the flat buffers (views stay views)."""
        super().load_state_dict(state_dict)
        for gi, group in enumerate(self.param_groups):
            self._flat.pop(gi, None)
            st = self._group_state(gi, group)
            steps = []
            for p in group["params"]:
                loaded = self.state.get(p)
                if not loaded or "exp_avg" not in loaded:
                    continue
                m, v, step = loaded["exp_avg"], loaded["exp_avg_sq"], int(loaded["step"])
                self.state[p] = {}
                s = self._param_state(st, p)
                s["exp_avg"].copy_(m.to(s["exp_avg"].dtype))
                s["exp_avg_sq"].copy_(v.to(s["exp_avg_sq"].dtype))
                s["step"] = step
                steps.append(step)
            if st["step_dev"] is not None and steps:
                st["step_dev"].fill_(max(steps))

    # ---- the step ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _capi.lib()
        for gi, group in enumerate(self.param_groups):
            st = self._group_state(gi, group)
            dev = st["device"]
            handle = _capi.context(dev.index if dev.index is not None else torch.cuda.current_device())
            active = [p for p in st["f32"] if p.grad is not None]
            a64 = st["f64"] if st["f64"] is not None and st["f64"].grad is not None else None
            if not active and a64 is None:
                continue
            for p in active + ([a64] if a64 is not None else []):
                if p.grad.is_sparse or not p.grad.is_contiguous() or p.grad.dtype != p.dtype or not p.is_contiguous():
                    raise RuntimeError("FusedAdam needs dense, contiguous gradients and parameters of the parameter's dtype")
                self._param_state(st, p)["step"] += 1
            capturing = torch.cuda.is_current_stream_capturing()
            if self.capturable:
                if not capturing:
                    st["lr_dev"].fill_(float(group["lr"]))
                by_step = {0: (active, a64)}      # one device counter for the whole group
            else:
                if capturing:
                    raise RuntimeError("FusedAdam(capturable=False) cannot be captured in a CUDA graph: the step number is a host value")
                by_step: Dict[int, tuple] = {}
                for p in active:
                    by_step.setdefault(self.state[p]["step"], ([], None))[0].append(p)
                if a64 is not None:
                    k = self.state[a64]["step"]
                    by_step[k] = (by_step.get(k, ([], None))[0], a64)
            beta1, beta2 = group["betas"]
            stream = torch.cuda.current_stream(dev).cuda_stream
            for step_no, (tensors, p64) in by_step.items():
                chunks: List[list] = [tensors[i:i + _MAX_TENSORS] for i in range(0, len(tensors), _MAX_TENSORS)] or [[]]
                for ci, chunk in enumerate(chunks):
                    n = len(chunk)
                    hyper = _capi.AdamHyper(float(group["lr"]), float(beta1), float(beta2), float(group["eps"]),
                                            float(group["weight_decay"]), max(int(step_no), 1))
                    # the pointer tables only change when a tensor moves (the fused backward keeps every gradient a view of one
                    # persistent flat buffer): build them once per layout
                    key = tuple(p.data_ptr() for p in chunk) + tuple(p.grad.data_ptr() for p in chunk)
                    cached = st.setdefault("tables", {}).get(key)
                    if cached is None:
                        if len(st["tables"]) > 8:
                            st["tables"].clear()
                        cached = ((C.c_void_p * max(n, 1))(*key[:n]), (C.c_void_p * max(n, 1))(*key[n:]),
                                  (C.c_int64 * max(n, 1))(*[p.numel() for p in chunk]),
                                  (C.c_int64 * max(n, 1))(*[st["offsets"][p] for p in chunk]))
                        st["tables"][key] = cached
                    pp, gp, ne, of = cached
                    last = ci == len(chunks) - 1
                    use64 = p64 is not None and last
                    # the device step counter ticks once per step: on the first launch of the group
                    step_dev = st["step_dev"].data_ptr() if (self.capturable and ci == 0) else None
                    if self.capturable and ci > 0:
                        raise RuntimeError("FusedAdam(capturable=True) supports at most 96 float32 tensors per parameter group")
                    with torch.cuda.device(dev):
                        rc = lib.dta_adam_step(handle, n, pp, gp, ne, of, st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                               p64.data_ptr() if use64 else None, p64.grad.data_ptr() if use64 else None,
                                               st["moments64"].data_ptr() if use64 else None, C.byref(hyper), step_dev,
                                               st["lr_dev"].data_ptr() if self.capturable else None, stream)
                    _capi.check(handle, rc, "dta_adam_step")
        return loss
