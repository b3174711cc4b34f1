"""Template-bank residency inside the reference's own inference loop.

run_test.py caches the template FEATURES once per object (run_test.py:121-134) but then, for every detection
batch, copies its objects' banks out of that cache (`templates_data[key][obj_idx].contiguous()`, :161-162,
170 MB per detection at the native size), normalises the copy (model/picopose.py:99) and hands it to
`matching_templates`, which normalises it again.  `as_bank` replaces the cached tensor by a `BankHandle`: a
storage-less tensor subclass that keeps the banks prepared once (bf16, K-major, inverse norms) and lets
exactly the operations the reference applies on the way to `matching_templates` pass through as index
bookkeeping:

    templates_data['template_feature'] = as_bank(torch.stack(template_features))      # the one edited line
    ...
    inputs[key] = templates_data[key][obj_idx].contiguous()      # -> BankHandle carrying obj_idx, no copy
    feature_tem = F.normalize(end_points['template_feature'], dim=2)   # -> the same handle (norms are applied
    matching_templates(feature_tem, ...)                               #    in the contraction's epilogue)

Anything else done to a handle raises: it has no data to compute on.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

_PASS_THROUGH = {"contiguous", "detach", "float", "cuda", "clone", "requires_grad_"}
_METADATA = {"size", "dim", "numel", "stride", "storage_offset", "is_contiguous", "is_floating_point", "is_complex",
             "__len__", "element_size", "nelement", "ndimension", "is_cuda", "get_device", "data_ptr"}


class BankHandle(torch.Tensor):
    """(B, N, C, H, W)-shaped stand-in for `bank[index]`; `bank` is a TemplateBank (or anything with n_banks, n_views,
    C, H, W, device), `index` a 1-D integer tensor of bank ids per detection (None = all banks in order)."""

    @staticmethod
    def __new__(cls, bank, index: Optional[torch.Tensor] = None):
        n = bank.n_banks if index is None else int(index.numel())
        r = torch.Tensor._make_wrapper_subclass(cls, (n, bank.n_views, bank.C, bank.H, bank.W), dtype=torch.float32,
                                                device=bank.device, requires_grad=False)
        r._bank, r._index = bank, index
        return r

    @property
    def bank(self):
        return self._bank

    @property
    def bank_index(self) -> Optional[torch.Tensor]:
        return self._index

    def _select(self, idx) -> "BankHandle":
        if isinstance(idx, tuple):
            if len(idx) != 1:
                raise NotImplementedError("BankHandle can only be indexed along its first (object) axis")
            idx = idx[0]
        base = self._index if self._index is not None else torch.arange(self._bank.n_banks, device=self._bank.device)
        if isinstance(idx, torch.Tensor):
            if idx.dtype == torch.bool or idx.dim() > 1:
                raise NotImplementedError("BankHandle takes a 1-D integer index (the detections' obj_idx)")
            picked = base[idx.to(base.device).long().reshape(-1)]
        elif isinstance(idx, (slice, list)):
            picked = base[idx].reshape(-1)
        else:
            raise NotImplementedError(f"BankHandle cannot be indexed with {type(idx).__name__}")
        return BankHandle(self._bank, picked)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        name = getattr(func, "__name__", str(func))
        self = args[0] if args else None
        if isinstance(self, BankHandle):
            if func is torch.Tensor.__getitem__:
                return self._select(args[1])
            if name in _PASS_THROUGH:
                return self
            if name == "to":
                return self          # device / dtype moves: the prepared bank stays where it is
            if func is F.normalize:
                dim = kwargs.get("dim", args[2] if len(args) > 2 else 1)
                p = kwargs.get("p", args[1] if len(args) > 1 else 2.0)
                if dim in (2, -3) and float(p) == 2.0:
                    return self      # channel normalisation is applied by the contraction's epilogue
                raise NotImplementedError("BankHandle only passes F.normalize(x, dim=2) (model/picopose.py:99) through")
            if name in _METADATA or name in ("__get__", "__repr__", "__format__", "__str__"):
                with torch._C.DisableTorchFunctionSubclass():
                    return func(*args, **kwargs)
        raise NotImplementedError(
            f"BankHandle stands for prepared template banks and holds no fp32 data: `{name}` is not supported "
            f"(supported: handle[obj_idx], .contiguous(), .to()/.cuda(), F.normalize(handle, dim=2), matching_templates)")

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        raise NotImplementedError(
            f"BankHandle holds no fp32 data: `{func}` is not supported (pass it to picopose_b200.matching.matching_templates)")

    def __repr__(self):
        idx = "all" if self._index is None else self._index.tolist()
        return f"BankHandle(shape={tuple(self.shape)}, banks={self._bank.n_banks}, index={idx})"


def as_bank(template_feature: torch.Tensor, mode: Optional[str] = None) -> BankHandle:
    """(n_obj, N, C, H, W) cached template features -> resident prepared banks behind a BankHandle."""
    from .matching import TemplateBank
    if isinstance(template_feature, BankHandle):
        return template_feature
    return BankHandle(TemplateBank.from_features(template_feature, mode))
