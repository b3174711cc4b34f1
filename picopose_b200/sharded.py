"""Multi-GPU stage-1 matching: template-axis sharding + one top-k exchange.

One process per GPU (torchrun).  Rank g keeps views [lo_g, hi_g) of every object's template bank
(prepared bf16, resident), all ranks see the whole detection batch (a rank's own detections come from its
host, the others' over NVLink: `ShardedMatcher.gather_queries`), each computes its local
sim_avg[:, lo_g:hi_g] and a local top-k with GLOBAL view indices, and the (B, k) candidate pairs are exchanged and merged
identically on every rank.  Every (detection, view) score is independent (utils/matching.py:47-67); only topk (:68)
couples views, hence one exchange.  Two forms of it:

* `PeerExchange` (default on CUDA): ONE kernel does local top-k, stores the pairs into every peer's buffer through the
  NVLink peer mapping (CUDA IPC), flags them, waits for the peers' flags and merges (csrc/exchange.cu);
* `merge_topk`: local top-k kernel, one NCCL (or gloo) all-gather of a (B, k, 2) tensor, merge kernel -- the portable
  form, used when peer mappings are not available and by the CPU tests of the plumbing.
"""
from __future__ import annotations

import os
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_views: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, as-even-as-possible split: the first n_views % world ranks get one extra view."""
    base, extra = divmod(n_views, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def topk_pairs(sim: torch.Tensor, k: int, idx_offset: int = 0) -> torch.Tensor:
    """Local top-k of (B, N_local) scores as one (B, k, 2) float64 tensor of (score, global index) pairs,
    padded with (-inf, -1) when the shard has fewer than k views."""
    from . import _lib
    _lib.require_cuda(sim)
    lib = _lib.load()
    sim = sim.float().contiguous()
    B, N = sim.shape
    out = torch.empty(B, k, 2, dtype=torch.float64, device=sim.device)
    if N == 0:
        out[..., 0] = float("-inf")
        out[..., 1] = -1
        return out
    with torch.cuda.device(sim.device):
        _lib.check(lib.pp_topk_pairs(_lib.ptr(sim), B, N, k, idx_offset, _lib.ptr(out), _lib.stream_of(sim)),
                   "pp_topk_pairs")
    return out


def _merge_cuda(gathered: torch.Tensor, k: int):
    """gathered (R, B, k_in, 2) float64 on a CUDA device -> (score (B,k) f32, idx (B,k) i64)."""
    from . import _lib
    lib = _lib.load()
    R, B, k_in, _ = gathered.shape
    score = torch.empty(B, k, dtype=torch.float32, device=gathered.device)
    idx = torch.empty(B, k, dtype=torch.int64, device=gathered.device)
    with torch.cuda.device(gathered.device):
        _lib.check(lib.pp_topk_merge(_lib.ptr(gathered), R, B, k_in, k, _lib.ptr(score), _lib.ptr(idx),
                                     _lib.stream_of(gathered)), "pp_topk_merge")
    return score, idx


def _merge_torch(gathered: torch.Tensor, k: int):
    """Same selection rule in plain torch ops; used by the CPU/gloo tests of the exchange plumbing only."""
    R, B, k_in, _ = gathered.shape
    flat = gathered.permute(1, 0, 2, 3).reshape(B, R * k_in, 2)
    order = torch.argsort(flat[..., 0], dim=1, descending=True, stable=True)[:, :k]   # stable: lowest rank/slot first
    picked = torch.gather(flat, 1, order.unsqueeze(-1).expand(B, k, 2))
    return picked[..., 0].float(), picked[..., 1].long()


def merge_topk(local_pairs: torch.Tensor, k: int, group=None, merge: Optional[Callable] = None):
    """One all-gather of every rank's (B, k, 2) candidate pairs, then the same merge on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    B, k_in, _ = local_pairs.shape
    local_pairs = local_pairs.contiguous()
    if world == 1:
        gathered = local_pairs.unsqueeze(0)
    else:
        # rank-major concatenation along dim 0 (the layout both NCCL and gloo accept)
        flat = torch.empty(world * B, k_in, 2, dtype=local_pairs.dtype, device=local_pairs.device)
        dist.all_gather_into_tensor(flat, local_pairs, group=group)
        gathered = flat.view(world, B, k_in, 2)
    merge = merge or (_merge_cuda if local_pairs.is_cuda else _merge_torch)
    return merge(gathered, k)


class PeerExchange:
    """Per-rank exchange buffer + peer mappings for `pp_topk_exchange` (one node, one process per GPU).

    Collective constructor: every rank allocates its buffer in the library (cudaMalloc + cudaIpcGetMemHandle), the
    64-byte handles travel through one `all_gather_object`, every rank maps its peers' buffers.  `exchange` is a
    collective too (every rank must call it, with the same B and k); the epoch counter advances identically everywhere.
    """

    def __init__(self, group=None, max_b: int = 64, k_max: int = 8, device=None):
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.max_b, self.k_max = int(max_b), int(k_max)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.epoch = 0
        self._own = None
        self._opened = []
        nbytes = lib.pp_xchg_bytes(self.world, self.max_b, self.k_max)
        self._own, ptrs, self._opened = _open_peers(lib, group, nbytes, self.device)
        self.peers = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
        dist.barrier(group)   # nobody pushes before everybody has mapped everybody

    def exchange(self, sim: torch.Tensor, k: int, idx_offset: int = 0):
        """sim (B, N_local) local scores -> global (score (B,k) f32, idx (B,k) i64), identical on every rank."""
        from . import _lib
        lib = _lib.load()
        sim = sim.float().contiguous()
        B, N = sim.shape
        if B > self.max_b or k > self.k_max:
            raise ValueError(f"PeerExchange sized for B <= {self.max_b}, k <= {self.k_max} (got B={B}, k={k})")
        self.epoch = 1 if self.epoch >= 0xFFFFFFFE else self.epoch + 1      # never 0, parity alternates across the wrap
        score = torch.empty(B, k, dtype=torch.float32, device=sim.device)
        idx = torch.empty(B, k, dtype=torch.int64, device=sim.device)
        with torch.cuda.device(sim.device):
            _lib.check(lib.pp_topk_exchange(_lib.ptr(sim), B, N, k, idx_offset, _lib.ptr(self.peers), self.rank, self.world,
                                            self.max_b, self.k_max, self.epoch, _lib.ptr(score), _lib.ptr(idx),
                                            _lib.stream_of(sim)), "pp_topk_exchange")
        return score, idx

    def close(self):
        """Collective: unmap the peers' buffers, then free the own one."""
        from . import _lib
        _close_peers(_lib.load(), self.group, self.device, self._own, self._opened)
        self._own, self._opened = None, []


class _DeviceView:
    """Zero-copy torch view of library-owned device memory (through __cuda_array_interface__)."""

    def __init__(self, ptr: int, shape, typestr: str = "<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def _open_peers(lib, group, nbytes: int, device):
    """Collective: allocate this rank's exchange buffer, swap the IPC handles, map every peer's buffer.
    -> (own pointer, [pointer of rank 0 .. world-1] with the own one in place, [opened peer pointers]).

    Failure is collective too: a rank whose allocation or mapping fails (out of memory, no P2P path to one peer) still
    takes part in both object all-gathers, every rank learns that somebody failed, releases what it had created and
    raises the same RuntimeError -- so the callers' fallback to the all-gather forms is taken by all ranks or by none,
    and nobody is left waiting inside a collective."""
    import ctypes as C
    from . import _lib
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    own, opened, ptrs, err = None, [], [], None
    with torch.cuda.device(device):
        buf, handle = C.c_void_p(), C.create_string_buffer(64)
        try:
            _lib.check(lib.pp_xchg_create(nbytes, C.byref(buf), handle), "pp_xchg_create")
            own = buf.value
        except RuntimeError as exc:
            err = str(exc)
        handles = [None] * world
        dist.all_gather_object(handles, handle.raw if own is not None else None, group=group)
        if err is None and all(h is not None for h in handles):
            try:
                for r, h in enumerate(handles):
                    if r == rank:
                        ptrs.append(own)
                    else:
                        p = C.c_void_p()
                        _lib.check(lib.pp_xchg_open(C.create_string_buffer(h, 64), C.byref(p)), "pp_xchg_open")
                        opened.append(p.value)
                        ptrs.append(p.value)
            except RuntimeError as exc:
                err = str(exc)
        elif err is None:
            err = "a peer could not allocate its exchange buffer"
        verdicts = [None] * world
        dist.all_gather_object(verdicts, err, group=group)
        failed = [(r, v) for r, v in enumerate(verdicts) if v is not None]
        if failed:
            for p in opened:
                lib.pp_xchg_close(p)
            dist.barrier(group)          # nobody frees a buffer a peer may still have mapped
            if own is not None:
                lib.pp_xchg_destroy(own)
            raise RuntimeError("peer-memory buffers unavailable on rank(s) %s: %s"
                               % ([r for r, _ in failed], failed[0][1]))
    return own, ptrs, opened


def _close_peers(lib, group, device, own, opened):
    """Collective: unmap the peers' buffers, then free the own one (nobody frees what a peer still has mapped)."""
    torch.cuda.synchronize(device)
    dist.barrier(group)
    with torch.cuda.device(device):
        for p in opened:
            lib.pp_xchg_close(p)
        dist.barrier(group)
        if own is not None:
            lib.pp_xchg_destroy(own)


class PeerGather:
    """All-gather of every rank's query slice (features + mask) through NVLink peer memory with the COPY ENGINES.

    An NCCL all-gather needs SMs, and the contraction kernel holds all of them, so a gather issued for step i+1 while
    step i computes only runs in the gaps between kernels.  Here rank r copies its slice into slot r of every peer's
    buffer with `cudaMemcpyAsync` peer copies (no SM), a one-block kernel then writes the call's epoch into flag [r]
    of every peer, and a one-block kernel spins on the `world` flags of the own buffer; all on the caller's stream.
    Slots are double-buffered by epoch parity.  Contract: (1) a `gather` may run at most one step ahead of the compute it
    feeds -- the stream it is issued on must be ordered after the `match` (top-k exchange) that consumed the gather before
    the previous one, as any double-buffered upload loop is; that exchange proves every peer has passed the query
    prologue of that step.  (2) The gathered tensors are views of this rank's buffer and only the query prologue of the
    step's `match` may read them: a faster peer is free to overwrite that parity as soon as the exchange of the same
    step has completed.  Pass `out=` to `ShardedMatcher.gather_queries` for a private copy.
    """

    def __init__(self, tar_slice_shape, mask_slice_shape, group=None, device=None):
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.tar_shape, self.mask_shape = tuple(tar_slice_shape), tuple(mask_slice_shape)
        self.tar_bytes = 4 * int(torch.Size(self.tar_shape).numel())
        self.mask_bytes = 4 * int(torch.Size(self.mask_shape).numel())
        assert self.tar_bytes % 16 == 0 and self.mask_bytes % 16 == 0
        lay = self.layout(self.world, self.tar_bytes, self.mask_bytes)
        self.flag_bytes, self.tar_region, self.parity_bytes = lay["flag_bytes"], lay["tar_region"], lay["parity_bytes"]
        nbytes = lay["total"]
        self._own, ptrs, self._opened = _open_peers(lib, group, nbytes, self.device)
        self._peers_host = (C.c_void_p * self.world)(*ptrs)
        self.peers = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
        self._fused = os.environ.get("PICOPOSE_B200_FUSED_GATHER", "1") != "0"   # 0: separate push / push / signal calls
        self.epoch = 0
        self.consumed = 0      # gathers whose batch has been handed to a `match` (ShardedMatcher keeps it up to date)
        self._issued = 0
        dist.barrier(group)

    @staticmethod
    def layout(world: int, tar_bytes: int, mask_bytes: int) -> dict:
        """Byte layout of one rank's buffer: flags [2][world] u32 | parity 0: tar slots, mask slots | parity 1: ...;
        every region starts on a 256-byte boundary."""
        up = lambda x: 256 * ((x + 255) // 256)                                    # noqa: E731
        flag_bytes = up(2 * world * 4)
        tar_region = world * tar_bytes
        mask_off = up(tar_region)
        parity_bytes = mask_off + up(world * mask_bytes)
        return {"flag_bytes": flag_bytes, "tar_region": tar_region, "mask_off": mask_off, "parity_bytes": parity_bytes,
                "total": flag_bytes + 2 * parity_bytes}

    def _offsets(self, par: int):
        base = self.flag_bytes + par * self.parity_bytes
        return base, base + 256 * ((self.tar_region + 255) // 256)

    def gather(self, tar_local: torch.Tensor, mask_local: torch.Tensor):
        """-> (tar (world*b, ...), mask (world*b, ...)) views of this rank's buffer, complete in stream order on the
        current stream (and on any stream that waits for it); see the class contract for how long they stay valid."""
        from . import _lib
        lib = _lib.load()
        assert tuple(tar_local.shape) == self.tar_shape and tuple(mask_local.shape) == self.mask_shape
        if self._issued - self.consumed >= 2:
            # contract (1): gather e+2 overwrites the slots of gather e, which is only safe once the match that consumed
            # gather e has run its exchange -- a third gather without a match in between would hand peers torn data
            raise RuntimeError("PeerGather: two gathers are already waiting for their `match`; a gather may run at most "
                               "one step ahead of the compute it feeds (double-buffered slots)")
        self._issued += 1
        tar_local, mask_local = tar_local.float().contiguous(), mask_local.float().contiguous()
        self.epoch = 1 if self.epoch >= 0xFFFFFFFE else self.epoch + 1
        par = self.epoch & 1
        t_off, m_off = self._offsets(par)
        st = _lib.stream_of(tar_local)
        with torch.cuda.device(self.device):
            if self._fused:
                # one library call for both pushes and the flag kernel (host time per gather 39 -> 22-30 us at 8 ranks)
                _lib.check(lib.pp_xchg_push_signal(_lib.ptr(tar_local), self.tar_bytes, t_off + self.rank * self.tar_bytes,
                                                   _lib.ptr(mask_local), self.mask_bytes, m_off + self.rank * self.mask_bytes,
                                                   self._peers_host, _lib.ptr(self.peers), par * self.world * 4, self.rank,
                                                   self.world, self.epoch, st), "pp_xchg_push_signal")
            else:
                _lib.check(lib.pp_xchg_push(_lib.ptr(tar_local), self.tar_bytes, self._peers_host, self.world,
                                            t_off + self.rank * self.tar_bytes, st), "pp_xchg_push")
                _lib.check(lib.pp_xchg_push(_lib.ptr(mask_local), self.mask_bytes, self._peers_host, self.world,
                                            m_off + self.rank * self.mask_bytes, st), "pp_xchg_push")
                _lib.check(lib.pp_xchg_signal(_lib.ptr(self.peers), par * self.world * 4, self.rank, self.world, self.epoch,
                                              st), "pp_xchg_signal")
            _lib.check(lib.pp_xchg_wait(self._own, par * self.world * 4, self.world, self.epoch, st), "pp_xchg_wait")
        b = self.tar_shape[0]
        tar = torch.as_tensor(_DeviceView(self._own + t_off, (self.world * b,) + self.tar_shape[1:]), device=self.device)
        mask = torch.as_tensor(_DeviceView(self._own + m_off, (self.world * b,) + self.mask_shape[1:]), device=self.device)
        return tar, mask

    def close(self):
        """Collective; tensors returned by `gather` must not be used afterwards."""
        from . import _lib
        _close_peers(_lib.load(), self.group, self.device, self._own, self._opened)
        self._own, self._opened = None, []


class ShardedMatcher:
    """Template-sharded `matching_templates` over a process group."""

    def __init__(self, n_views: int, group=None, merge: Optional[Callable] = None, query_group=None,
                 peer_exchange: Optional[bool] = None):
        self.group = group
        # gather_queries on its own communicator: collectives of one NCCL communicator run in issue order on one
        # internal stream, so a query gather for step i+1 would otherwise queue behind step i's top-k exchange
        self.query_group = query_group if query_group is not None else group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_views = n_views
        self.lo, self.hi = shard_range(n_views, self.rank, self.world)
        self.merge = merge
        self.bank = None
        # top-k exchange through NVLink peer memory unless a custom merge was injected or it is switched off
        if peer_exchange is None:
            peer_exchange = merge is None and os.environ.get("PICOPOSE_B200_PEER_EXCHANGE", "1") != "0"
        self._want_peer = bool(peer_exchange) and self.world > 1
        self._xchg = None
        self._gather = None

    def load_bank(self, src_feats_shard: torch.Tensor, mode: Optional[str] = None):
        """src_feats_shard: this rank's views (n_banks, hi-lo, C, H, W) fp32 -> resident prepared bank."""
        from .matching import TemplateBank
        assert src_feats_shard.shape[1] == self.hi - self.lo
        self.bank = TemplateBank.from_features(src_feats_shard, mode)
        return self.bank

    def gather_queries(self, tar_local: torch.Tensor, mask_local: torch.Tensor, out=None):
        """Each rank received its own detections' query features (b, C, H, W) and masks (b, Hm, Wm) from its host;
        every rank needs the whole batch (template-axis sharding), so the queries travel GPU-to-GPU over NVLink
        instead of every rank uploading world x the data over PCIe: copy-engine peer pushes (`PeerGather`, see its
        one-step-ahead contract) or, as fallback, two NCCL all-gathers.
        -> (tar (world*b, C, H, W), mask (world*b, Hm, Wm)), rank-major; `out` = optional preallocated pair (filled
        and returned; leave it out to get zero-copy views of the exchange buffer, which only the following `match` may
        read -- see `PeerGather`)."""
        if self.world == 1:
            return tar_local, mask_local
        tar_local, mask_local = tar_local.contiguous(), mask_local.contiguous()
        if self._want_peer and tar_local.is_cuda and tar_local.dtype == torch.float32 and mask_local.dtype == torch.float32:
            # copy-engine pushes through peer memory (PeerGather): no SMs, so it overlaps the contraction of the step
            # before; without `out` the result is a view of this rank's exchange buffer
            if self._gather is None or self._gather.tar_shape != tuple(tar_local.shape) \
                    or self._gather.mask_shape != tuple(mask_local.shape):
                if self._gather is not None:
                    self._gather.close()
                    self._gather = None
                try:    # collective: fails on every rank or on none (_open_peers)
                    self._gather = PeerGather(tar_local.shape, mask_local.shape, self.group, tar_local.device)
                except RuntimeError as exc:
                    import warnings
                    warnings.warn(f"picopose_b200: peer-memory query gather unavailable ({exc}); using NCCL all-gathers")
                    self._want_peer = False
            if self._gather is not None:
                tar_all, mask_all = self._gather.gather(tar_local, mask_local)
                if out is None:
                    return tar_all, mask_all                       # zero-copy views of this rank's exchange buffer
                out[0].copy_(tar_all)                              # a caller-provided pair is still filled
                out[1].copy_(mask_all)
                return out
        if out is None:
            out = (tar_local.new_empty((self.world * tar_local.shape[0],) + tuple(tar_local.shape[1:])),
                   mask_local.new_empty((self.world * mask_local.shape[0],) + tuple(mask_local.shape[1:])))
        dist.all_gather_into_tensor(out[0], tar_local, group=self.query_group)
        dist.all_gather_into_tensor(out[1], mask_local, group=self.query_group)
        return out

    def match(self, src, tar_feat, tar_mask, topk=5, bank_index=None, mode=None):
        """src: this rank's shard, a TemplateBank or raw (B|n_banks, hi-lo, C, H, W) features."""
        from .matching import matching_templates, template_scores
        src = self.bank if src is None else src
        if self._gather is not None and self._gather.consumed < self._gather._issued:
            self._gather.consumed += 1       # this match consumes the oldest outstanding gather (see PeerGather contract)
        if self.world == 1:
            # single rank: nothing to exchange, one library call ranks the whole bank
            return matching_templates(src, tar_feat, None, tar_mask, topk, mode=mode, bank_index=bank_index)
        sim = template_scores(src, tar_feat, tar_mask, mode=mode, bank_index=bank_index)     # (B, hi-lo)
        xchg = self._peer_exchange_for(sim, topk)
        if xchg is not None:
            return xchg.exchange(sim, topk, idx_offset=self.lo)
        return merge_topk(topk_pairs(sim, topk, idx_offset=self.lo), topk, self.group, self.merge)

    @property
    def uses_peer_memory(self) -> bool:
        """True while the exchange / gather go through NVLink peer memory (False: NCCL all-gather forms)."""
        return self._want_peer

    def close(self):
        """Collective: releases the peer-memory buffers (exchange + gather) if any were created."""
        for obj in (self._xchg, self._gather):
            if obj is not None:
                obj.close()
        self._xchg = self._gather = None

    def _peer_exchange_for(self, sim: torch.Tensor, k: int):
        """The PeerExchange to use for this call, created (collectively) on first use; None = all-gather form."""
        if not self._want_peer or not sim.is_cuda:
            return None
        B = sim.shape[0]
        if self._xchg is not None and (B > self._xchg.max_b or k > self._xchg.k_max):
            self._xchg.close()
            self._xchg = None
        if self._xchg is None:
            try:
                self._xchg = PeerExchange(self.group, max_b=max(64, B), k_max=max(8, k), device=sim.device)
            except RuntimeError as exc:   # collective failure (_open_peers): every rank lands here together
                import warnings
                warnings.warn(f"picopose_b200: peer-memory top-k exchange unavailable ({exc}); using the all-gather form")
                self._want_peer = False
                return None
        return self._xchg
