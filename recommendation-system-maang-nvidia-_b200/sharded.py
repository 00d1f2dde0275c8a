"""Catalog partitioned over the GPUs of one node (BASELINE config 3).

One process per GPU.  Rank g holds the contiguous catalog rows [bounds[g], bounds[g+1]) in its own
IndexFlatIP; queries are replicated.  A search is: local scan + local top-k (global row numbers via
id_base) -> ONE all-gather of each shard's [nq, k] scores, rows and per-query flags over NCCL/NVLink
(G*nq*(12k+4) bytes, ~10 MB at G=8, nq=1024, k=100) -> on-device G-way merge (K7) straight out of the exchange buffer.  The tie-break (score desc, row asc) makes the merge
independent of the shard order, so every rank ends with the same lists as a single-GPU search over the
whole catalog -- bit for bit.

torch.distributed is the plumbing (process group, all_gather_into_tensor); the scan and the merge are this
repo's kernels -- and with exchange = "p2p" so is the exchange itself: PeerExchange (csrc/peer.cu) stores every rank's
message straight into its peers' memory over NVLink, and the merge reads the parts in place.  `local_search` / `merge` are injectable so the host logic (bounds, id offsets, gather
layout) can be exercised on CPU with gloo, where the tests plug the oracle in as the checker.
"""

from typing import Callable, List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_total: int, world: int) -> List[int]:
    """Contiguous row ranges: shard g = [g*N//G, (g+1)*N//G)."""
    return [n_total * g // world for g in range(world + 1)]


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def exchange_layout(nq: int, k: int) -> Tuple[int, int, int]:
    """Byte offsets of one shard's message in the fused exchange: scores f32 [nq,k] at 0, rows int64 [nq,k] at
    off_I, need_exact int32 [nq] at off_F, then bound f32 [nq] (shared-threshold searches) at off_F + 4*nq;
    `total` bytes per shard (16-byte multiples)."""
    off_I = _round_up(nq * k * 4, 16)
    off_F = off_I + nq * k * 8
    return off_I, off_F, _round_up(off_F + nq * 8, 16)


class _DeviceBytes:
    """Raw device memory as a __cuda_array_interface__ object, so torch can view it without a copy."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class PeerExchange:
    """All-gather of one byte message per rank over NVLink peer memory (csrc/peer.cu) instead of NCCL.

    Every rank owns a region of `world` slots, exported through CUDA IPC; `publish(msg)` copies the message into slot
    `rank` of every rank's region and raises a flag there, `gathered()` returns -- stream-ordered behind a kernel that
    waits for all flags -- the local region as a uint8 [world, slot_bytes] tensor (no copy: the merge kernel reads
    the parts in place), `done()` tells the peers the region may be overwritten by the next step.  Construction is
    collective (the IPC handles are all-gathered once through torch.distributed); the three per-step calls are plain
    kernel launches and can be captured in a CUDA graph."""

    def __init__(self, slot_bytes: int, device: torch.device, group: Optional[dist.ProcessGroup] = None):
        import ctypes
        from . import _lib
        self._lib_mod, self._lib = _lib, _lib.load()
        multi = dist.is_initialized()
        self.world, self.rank = (dist.get_world_size(group), dist.get_rank(group)) if multi else (1, 0)
        self.slot_bytes = _round_up(int(slot_bytes), 16)
        self.device = device
        self._group = group
        self._h = ctypes.c_void_p()
        # Every rank walks through the same collectives whether or not its own CUDA calls succeed; the outcome is
        # agreed on (MIN over ranks) and a failure anywhere raises RuntimeError on EVERY rank.
        handle = (ctypes.c_uint8 * 64)()
        ok, err = 1, ""
        try:
            _lib.check(self._lib.rs_peer_create(self.rank, self.world, self.slot_bytes, ctypes.byref(self._h), handle))
        except RuntimeError as e:
            ok, err = 0, str(e)
        blob = bytes(handle)
        if multi:
            mine = torch.tensor(list(handle), dtype=torch.uint8).to(device)
            everyone = torch.empty((self.world * 64,), dtype=torch.uint8, device=device)
            dist.all_gather_into_tensor(everyone, mine, group=group)
            blob = bytes(everyone.cpu().tolist())
        if ok:
            try:
                _lib.check(self._lib.rs_peer_connect(self._h, ctypes.c_char_p(blob)))
            except RuntimeError as e:
                ok, err = 0, str(e)
        if multi:
            flag = torch.tensor([ok], dtype=torch.int32, device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            ok_all = int(flag.item())
        else:
            ok_all = ok
        if not ok_all:
            self.region = None
            self.close()
            raise RuntimeError("peer-memory region could not be set up on every rank" + (f": {err}" if err else ""))
        ptr, n = ctypes.c_void_p(), ctypes.c_int64()
        _lib.check(self._lib.rs_peer_local(self._h, ctypes.byref(ptr), ctypes.byref(n)))
        self._mem = _DeviceBytes(ptr.value, n.value)
        self.region = torch.as_tensor(self._mem, device=device).view(self.world, self.slot_bytes)
        if multi:
            dist.barrier(group=group)      # nobody publishes into a region that is not mapped everywhere yet

    def publish(self, msg: torch.Tensor) -> None:
        b = self._lib_mod.dl(msg)
        self._lib_mod.check(self._lib.rs_peer_publish(self._h, b.ptr, self._lib_mod.stream_ptr()))

    def gathered(self) -> torch.Tensor:
        self._lib_mod.check(self._lib.rs_peer_wait(self._h, self._lib_mod.stream_ptr()))
        return self.region

    def done(self) -> None:
        self._lib_mod.check(self._lib.rs_peer_ack(self._h, self._lib_mod.stream_ptr()))

    def close(self) -> None:
        """Collective: every rank of the group calls it.  Nobody frees a region that a peer may still write into or
        still has mapped: finish all work -> barrier -> unmap the peers' regions -> barrier -> free the own one."""
        if getattr(self, "_h", None) is None:
            return
        multi = self.world > 1 and dist.is_initialized()
        torch.cuda.synchronize(self.device)
        if multi:
            dist.barrier(group=self._group)
        self.region = None
        if self._h.value:
            self._lib_mod.check(self._lib.rs_peer_disconnect(self._h))
        if multi:
            dist.barrier(group=self._group)
        if self._h.value:
            self._lib_mod.check(self._lib.rs_peer_free(self._h))
        self._h = None


class Comm:
    """rs_comm_t (include/recsys_b200.h) over a torch.distributed process group: this rank's view of the group plus the
    host all-gather callback the library uses to set up and tear down its peer-memory regions (CUDA IPC handles,
    barriers).  Per-step traffic never goes through it.  One communicator per (group, device), shared by the sharded
    index and the sharded tables; construction, close() and everything created on it are collective."""

    _live = {}

    @classmethod
    def get(cls, device: torch.device, group: Optional[dist.ProcessGroup] = None, lane: int = 0) -> "Comm":
        key = (id(group) if group is not None else 0, torch.device(device).index, int(lane))
        c = cls._live.get(key)
        if c is None or c._h is None:
            c = cls._live[key] = cls(device, group)
        c._refs += 1
        return c

    def release(self) -> None:
        """Drop one user; the last one frees the communicator (collective: every rank releases in the same order)."""
        self._refs -= 1
        if self._refs <= 0:
            self.close()

    def __init__(self, device: torch.device, group: Optional[dist.ProcessGroup] = None):
        import ctypes
        from . import _lib
        self._lib_mod, self._lib = _lib, _lib.load()
        multi = dist.is_initialized()
        self.world, self.rank = (dist.get_world_size(group), dist.get_rank(group)) if multi else (1, 0)
        self.group, self.device = group, torch.device(device)
        on_dev = multi and dist.get_backend(group) == "nccl"
        world = self.world

        def allgather(ctx, send, recv, n):
            # called by the library with host pointers; every rank calls it at the same point of the same C function
            try:
                n = int(n)
                mine = torch.frombuffer(bytearray(ctypes.string_at(send, n)), dtype=torch.uint8)
                if on_dev:
                    mine = mine.to(self.device)
                everyone = torch.empty((world * n,), dtype=torch.uint8, device=mine.device)
                dist.all_gather_into_tensor(everyone, mine, group=group)
                host = everyone.cpu().contiguous()
                ctypes.memmove(recv, host.data_ptr(), world * n)
                return 0
            except Exception:           # must not propagate through the C frames: the library reports the failure
                import traceback
                traceback.print_exc()
                return 1

        self._refs = 0
        self._cb = _lib.ALLGATHER_FN(allgather)          # kept alive as long as the communicator
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.rs_comm_init(self.rank, self.world, self._cb, None, ctypes.byref(self._h)))

    def close(self) -> None:
        if getattr(self, "_h", None) is None:
            return
        h, self._h = self._h, None
        with torch.cuda.device(self.device):
            self._lib_mod.check(self._lib.rs_comm_free(h))


class ShardedIndexFlatIP:
    def __init__(self, d: int, n_total: int, group: Optional[dist.ProcessGroup] = None,
                 local_index=None, local_search: Optional[Callable] = None, merge: Optional[Callable] = None,
                 local_deferred: Optional[Callable] = None, shared: Optional[Tuple[Callable, Callable]] = None,
                 plan: Optional[Callable] = None, rank: Optional[Callable] = None, verify: Optional[Callable] = None):
        self.d = int(d)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.bounds = shard_bounds(int(n_total), self.world)
        self.row_begin, self.row_end = self.bounds[self.rank], self.bounds[self.rank + 1]
        self.n_total = int(n_total)
        if self.n_total > 0xFFFFFFFF:
            # the merge packs (score, global row) into one 64-bit key with a 32-bit row field (csrc/flat_index.cu)
            raise ValueError("sharded catalogs above 2^32 rows are not supported (global row numbers are merged as 32-bit keys)")
        self.fallbacks = 0          # queries re-searched exactly after the exchange (results identical either way)
        if local_search is None:
            from .index import IndexFlatIP
            from . import ops
            self.local = local_index if local_index is not None else IndexFlatIP(d)
            self._local_search = lambda q, k: self.local.search(q, k, id_base=self.row_begin)
            self._local_exact = lambda q, k: self.local.search(q, k, mode="exact", id_base=self.row_begin)
            self._local_deferred = lambda q, k, out: self.local.search_deferred(q, k, id_base=self.row_begin, out=out)
            self._merge = ops.topk_merge
            self._verify = ops.topk_verify
            self._shared = (self.local.shared_begin,
                            lambda sess, tops, rank, out: self.local.shared_finish(sess, tops, rank, id_base=self.row_begin, out=out))
            self._plan = lambda rows, nq, k: IndexFlatIP.tensor_plan(self.d, rows, nq, k)
            self._rank = IndexFlatIP.threshold_rank
        else:
            self.local = local_index
            self._local_search = local_search
            self._local_exact = local_search
            self._local_deferred = local_deferred
            self._merge = merge
            self._verify = verify
            self._shared = shared        # (begin(q, k, top) -> session, finish(session, tops, rank, out))
            self._plan = plan            # (shard rows, nq, k) -> (tensor path applies, rows its sample covers)
            self._rank = rank            # (k, rho) -> rank of the sample maximum to use as the threshold
        self.shared_threshold = True     # set False to keep per-shard thresholds (one exchange fewer, ~8x the candidates)
        self.use_graph = False           # replay each (nq, k) step as one CUDA graph (CUDA queries only)
        self.reuse_outputs = False       # graph mode: return graph-owned buffers (two alternating graphs) instead of copies
        self.lane = 0                    # lanes of one index (submit / finish on separate streams) use separate peer regions
        self._flip = {}
        self.exchange = "nccl"           # "p2p": both exchanges of a step through NVLink peer memory (PeerExchange)
        self._peers = {}
        self._peer_factory = None        # tests: nbytes -> object with publish / gathered / done / close
        self.native = local_search is None   # exchange == "p2p": one rs_index_search_sharded call per step (csrc/sharded.cu)
        self._comm = None
        self._graphs = {}
        self._last_need = None

    @property
    def ntotal(self) -> int:
        return self.n_total

    def add_local(self, x) -> None:
        """Append rows of this rank's range (callers feed rows row_begin.. in order)."""
        self.local.add(x)
        if self.local.ntotal > self.row_end - self.row_begin:
            raise ValueError("more rows added than this rank's shard holds")

    def _shared_rank(self, nq: int, k: int) -> int:
        """Rank of the sample maximum that serves as the threshold shared by all shards; 0 = not applicable (some shard
        is outside the tensor path for this (nq, k), or the union of the top lists would not fit the kernel)."""
        key = (nq, k)
        cache = self.__dict__.setdefault("_rank_cache", {})
        if key not in cache:
            sampled, ok = 0, True
            for g in range(self.world):
                fits, rows = self._plan(self.bounds[g + 1] - self.bounds[g], nq, k)
                ok = ok and fits
                sampled += rows
            r = self._rank(k, sampled / max(1, self.n_total)) if ok else 0
            # every shard sends its r largest sample maxima; its sample has at least 256 of them
            cache[key] = r if (ok and r <= 128 and r * self.world <= 1024) else 0
        return cache[key]

    def _gather_merge(self, Dl: torch.Tensor, Il: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        nq, kk = int(Dl.shape[0]), int(Dl.shape[1])
        # rank-major concatenation along dim 0 (the layout both NCCL and gloo accept), viewed as [G, nq, k]
        Dg = torch.empty((self.world * nq, kk), dtype=Dl.dtype, device=Dl.device)
        Ig = torch.empty((self.world * nq, kk), dtype=Il.dtype, device=Il.device)
        dist.all_gather_into_tensor(Dg, Dl.contiguous(), group=self.group)
        dist.all_gather_into_tensor(Ig, Il.contiguous(), group=self.group)
        return self._merge(Dg.view(self.world, nq, kk), Ig.view(self.world, nq, kk))

    def search(self, q, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """q: [nq, d] replicated on every rank -> (D [nq, k], I [nq, k] global rows), identical on every rank.

        With a deferred local search (the default on CUDA) a step is queued without any host synchronisation in the
        middle: local search -> ONE all-gather of a byte buffer that carries each shard's scores, rows, need_exact
        flags and bounds side by side -> merge straight out of that buffer.  Only then does the host look at the flags
        (the step's one synchronisation); flagged queries -- the tensor path's rare unverified guesses -- are
        re-searched exactly on every shard and merged again.

        shared_threshold (default): the tensor path's filter threshold is guessed once from the union of all shards'
        samples (a small extra all-gather of [nq, rank] floats) instead of once per shard, so the shards together keep
        as many candidates as one index over the whole catalog would (~6.6 per cent of that each at G = 8 ... the
        filter's survivor path and the re-score shrink with it), and the guarantee moves behind the merge
        (ops.topk_verify: the k-th merged score must clear every shard's tau + eps)."""
        return self.finish(self.submit(q, k))

    def submit(self, q, k: int):
        """Queue one search step on the current stream WITHOUT looking at its flags -> a ticket for finish().  Lets a
        caller keep two batches in flight (two ShardedIndexFlatIP lanes over the same local index, each on its own
        stream): the latency-bound selection / exchange / merge kernels of one batch then overlap the tensor-bound filter
        of the other.  In graph mode with reuse_outputs the ticket's buffers are graph-owned and stay valid until the
        second-next submit on this lane."""
        k = int(k)
        cuda = isinstance(q, torch.Tensor) and q.is_cuda
        if self.world == 1 and not (self.use_graph and self._local_deferred is not None and cuda):
            return ("done",) + tuple(self._local_search(q, k))
        if self.world > 1 and self._local_deferred is None:
            return ("done",) + tuple(self._gather_merge(*self._local_search(q, k)))
        if self.use_graph and cuda:
            # graph mode (also on one GPU: the plain call synchronises with the host inside every search and pays its ~10
            # launches and stream-ordered allocations on the host each time; replayed, a step costs one graph launch)
            D, I, need_any = self._replay(q, k)
        else:
            D, I, need = self._step(q, k)
            need_any = need.any()
        return ("open", D, I, need_any, self._last_need, q, k)

    def finish(self, ticket) -> Tuple[torch.Tensor, torch.Tensor]:
        """The one host synchronisation of a step: read the flags; flagged queries -- the tensor path's rare unverified
        guesses, the same on every rank -- are re-searched exactly on every shard and merged again."""
        if ticket[0] == "done":
            return ticket[1], ticket[2]
        _, D, I, need_any, need, q, k = ticket
        if bool(need_any):
            idx = need.nonzero().flatten()
            sub = np.ascontiguousarray(q[idx.cpu().numpy()]) if isinstance(q, np.ndarray) else q[idx].contiguous()
            if self.world == 1:
                De, Ie = self._local_exact(sub, k)
            elif self._use_native(sub):
                De, Ie, _ = self._native_step(sub, int(k), mode="exact")    # the flags are the same on every rank
            else:
                De, Ie = self._gather_merge(*self._local_exact(sub, k))
            D[idx], I[idx] = De, Ie
            self.fallbacks += int(idx.numel())
        return D, I

    def _use_native(self, q) -> bool:
        return (self.native and self.exchange == "p2p" and self.shared_threshold and self._peer_factory is None
                and isinstance(q, torch.Tensor) and q.is_cuda)

    def _native_step(self, q: torch.Tensor, k: int, mode: str = "auto"):
        """One whole step as ONE C call (rs_index_search_sharded): sample, both peer-memory exchanges, filter, re-score,
        merge and verification are sequenced by the library on the current stream -> (D, I, need bool [nq])."""
        import ctypes
        from . import _lib
        if self._comm is None:
            self._comm = Comm.get(q.device, self.group, self.lane)
        nq = int(q.shape[0])
        D = torch.empty((nq, k), dtype=torch.float32, device=q.device)
        I = torch.empty((nq, k), dtype=torch.int64, device=q.device)
        need = torch.empty((nq,), dtype=torch.int32, device=q.device)
        bq, bD, bI, bN = _lib.dl(q.contiguous()), _lib.dl(D), _lib.dl(I), _lib.dl(need)
        _lib.check(_lib.load().rs_index_search_sharded(self.local._h, self._comm._h, self.n_total, bq.ptr, k, _lib.mode_code(mode),
                                                       bD.ptr, bI.ptr, bN.ptr, _lib.stream_ptr()))
        return D, I, need.ne(0)

    def close(self) -> None:
        """Drop the captured graphs (they hold NCCL work of the process group: release them before destroying it)."""
        if self._graphs:
            torch.cuda.synchronize()
            self._graphs.clear()
            torch.cuda.synchronize()
        for pair in self._peers.values():
            for pe in pair:
                if pe is not None:
                    pe.close()
        self._peers.clear()
        if self._comm is not None:
            torch.cuda.synchronize()
            self._comm.release()
            self._comm = None

    def _peer_pair(self, nq: int, k: int, rank_pad: int, dev):
        """(tops exchange or None, lists exchange) for this (nq, k): created collectively on first use.  If any rank
        cannot map its peers (no CUDA IPC / peer access on this box) every rank falls back to the NCCL all-gathers."""
        key = (nq, k, rank_pad)
        if key not in self._peers:
            make = self._peer_factory or (lambda nbytes: PeerExchange(nbytes, dev, self.group))
            pair, ok = [None, None], 1
            try:
                if rank_pad > 0:
                    pair[0] = make(nq * rank_pad * 4)
                pair[1] = make(exchange_layout(nq, k)[2])
            except RuntimeError as e:          # librecsys_b200 reports CUDA errors as RuntimeError
                ok, self._peer_error = 0, str(e)
            flag = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            if int(flag.item()) == 0:
                for pe in pair:
                    if pe is not None:
                        pe.close()
                import warnings
                warnings.warn("peer-memory exchange unavailable on some rank (%s); using NCCL all-gathers"
                              % getattr(self, "_peer_error", "failure on another rank"))
                self.exchange = "nccl"
                return None
            self._peers[key] = tuple(pair)
        return self._peers[key]

    def _step(self, q, k: int):
        """Queues one whole search step (no host synchronisation) -> (D, I, need bool [nq])."""
        if self.world == 1:
            nq1 = int(q.shape[0])
            D = torch.empty((nq1, k), dtype=torch.float32, device=q.device)
            I = torch.empty((nq1, k), dtype=torch.int64, device=q.device)
            F = torch.zeros((nq1,), dtype=torch.int32, device=q.device)
            self._local_deferred(q, k, (D, I, F))
            need = F.ne(0)
            self._last_need = need
            return D, I, need
        if self._use_native(q):
            try:
                D, I, need = self._native_step(q, k)
                self._last_need = need
                return D, I, need
            except RuntimeError as e:
                # regions are set up collectively inside the call: a failure anywhere is raised on every rank
                if "could not be set up on every rank" not in str(e):
                    raise
                import warnings
                warnings.warn(f"peer-memory exchange unavailable ({e}); using NCCL all-gathers")
                self.exchange = "nccl"
        nq = int(q.shape[0])
        G = self.world
        off_I, off_F, total = exchange_layout(nq, k)
        dev = q.device if isinstance(q, torch.Tensor) else torch.device("cpu")
        send = torch.empty((total,), dtype=torch.uint8, device=dev)
        out = (send[:nq * k * 4].view(torch.float32).view(nq, k),
               send[off_I:off_I + nq * k * 8].view(torch.int64).view(nq, k),
               send[off_F:off_F + nq * 4].view(torch.int32),
               send[off_F + nq * 4:off_F + nq * 8].view(torch.float32))
        rank = self._shared_rank(nq, k) if (self.shared_threshold and self._shared is not None) else 0
        p2p = self.exchange == "p2p" and (dev.type == "cuda" or self._peer_factory is not None)
        pe_tops = pe_lists = None
        if p2p:
            # peer-memory messages are copied with 128-bit stores: the tops carry rank rounded up to 4 columns
            pair = self._peer_pair(nq, k, _round_up(rank, 4), dev)
            if pair is None:
                p2p = False
            else:
                pe_tops, pe_lists = pair
        rank_pad = _round_up(rank, 4) if p2p else rank
        if rank > 0:
            # threshold shared by all shards: each contributes the `rank` largest maxima of its own sample
            begin, finish = self._shared
            top = torch.empty((nq, rank_pad), dtype=torch.float32, device=dev)
            sess = begin(q, k, top)
            if p2p:
                pe_tops.publish(top.view(-1).view(torch.uint8))
                tops = pe_tops.gathered().view(torch.float32).view(G, nq, rank_pad)
                finish(sess, tops, rank, out)
                pe_tops.done()
            else:
                tops = torch.empty((G * nq, rank), dtype=torch.float32, device=dev)
                dist.all_gather_into_tensor(tops, top, group=self.group)
                finish(sess, tops.view(G, nq, rank), rank, out)
        else:
            self._local_deferred(q, k, out[:3])
        if p2p:
            pe_lists.publish(send)
            R = pe_lists.gathered()                                              # [G, total] in place, no copy
        else:
            recv = torch.empty((G * total,), dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(recv, send, group=self.group)
            R = recv.view(G, total)
        Dg = R[:, :nq * k * 4].view(torch.float32).view(G, nq, k)            # parts strided by `total` bytes
        Ig = R[:, off_I:off_I + nq * k * 8].view(torch.int64).view(G, nq, k)
        Fg = R[:, off_F:off_F + nq * 4].view(torch.int32)                    # [G, nq]
        D, I = self._merge(Dg, Ig)
        if rank > 0:
            Bg = R[:, off_F + nq * 4:off_F + nq * 8].view(torch.float32)     # [G, nq]: every shard's tau + eps
            need = self._verify(D, I, Bg, Fg).ne(0)
        else:
            need = Fg.ne(0).any(dim=0)                                       # the same on every rank
        if p2p:
            pe_lists.done()                                                  # merge / verify have read the region
        self._last_need = need
        return D, I, need

    def _replay(self, q: torch.Tensor, k: int):
        """The step as a CUDA graph per (nq, k): ~15 launches and two NCCL collectives cost one graph launch on the
        host, which is what bounds a step once the catalog is split 8 ways (0.37 ms of kernels per step).  The graph
        owns its buffers; results are copied out so that the next replay cannot overwrite what the caller holds."""
        # reuse_outputs: two graphs per shape, replayed alternately, hand out their own output buffers instead of copies --
        # a result then stays valid until the second-next search of the same shape (what a serving loop needs)
        slot = 0
        if self.reuse_outputs:
            slot = self._flip.get((int(q.shape[0]), k), 0)
            self._flip[(int(q.shape[0]), k)] = slot ^ 1
        key = (int(q.shape[0]), k, q.device.index, slot)
        entry = self._graphs.get(key)
        if entry is None:
            for _ in range(2):                     # warm-up outside capture: lazy one-time setup inside the library
                self._step(q, k)
            torch.cuda.synchronize(q.device)
            static_q = q.clone()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                D, I, need = self._step(static_q, k)
                flag = need.any()
            entry = self._graphs[key] = (graph, static_q, D, I, need, flag)
        graph, static_q, D, I, need, flag = entry
        static_q.copy_(q)
        graph.replay()
        self._last_need = need
        if self.reuse_outputs:
            return D, I, flag
        return D.clone(), I.clone(), flag


def gather_bytes(world: int, nq: int, k: int) -> int:
    """Bytes each rank receives in the exchange step (f32 score + i64 row per hit, int32 flag per query)."""
    return world * exchange_layout(nq, k)[2]


class ShardedEmbedding:
    """Embedding table sharded over the ranks of a process group (BASELINE config 4).

    Global row r lives on rank r % G at local row r // G (even load under skew); the batch is data-parallel.
    lookup(ids) on every rank:
        plan      K7b kernels: per-owner histogram, then scatter of (r // G) into per-owner send segments + the
                  permutation that restores the batch order
        exchange  all-to-all #1 (NCCL over NVLink): local row numbers to their owners
        gather    K1 kernel on the owner's shard
        exchange  all-to-all #2: the 512-byte rows back to the requesters
        restore   K1 kernel with the permutation as ids
    Bytes over NVLink per looked-up row: 8 (id) out, 4*E (row) back, for the (G-1)/G share that is remote.
    enable_peer_exchange(max_batch) replaces both all-to-alls by stores into NVLink peer memory (csrc/peer.cu): hash
    de-duplication of the batch, device-known counts, owner-side gather written straight into the requester's region,
    no host synchronisation -- see _lookup_p2p.
    `plan` / `gather` are injectable so the host logic can run on CPU over gloo with numpy stand-ins (tests).
    """

    def __init__(self, n_rows_total: int, dim: int, local_table=None, group: Optional[dist.ProcessGroup] = None,
                 plan: Optional[Callable] = None, gather: Optional[Callable] = None):
        self.n_rows_total, self.dim, self.group = int(n_rows_total), int(dim), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_local = (self.n_rows_total - self.rank + self.world - 1) // self.world
        self.local_table = local_table
        if plan is None:
            from . import _lib, ops

            def plan(ids, G):
                counts = torch.empty((G,), dtype=torch.int64, device=ids.device)
                a, b = _lib.dl(ids), _lib.dl(counts)
                _lib.check(_lib.load().rs_a2a_count(a.ptr, G, b.ptr, _lib.stream_ptr()))
                return counts

            def scatter(ids, G, base):
                cursor = torch.empty((G,), dtype=torch.int64, device=ids.device)
                send = torch.empty_like(ids)
                perm = torch.empty_like(ids)
                a, b, c, d, e = _lib.dl(ids), _lib.dl(base), _lib.dl(cursor), _lib.dl(send), _lib.dl(perm)
                _lib.check(_lib.load().rs_a2a_scatter(a.ptr, G, b.ptr, c.ptr, d.ptr, e.ptr, _lib.stream_ptr()))
                return send, perm

            self._count, self._scatter = plan, scatter
            self._gather = lambda table, rows: ops.gather_concat([table], [rows])
        else:
            self._count, self._scatter = plan
            self._gather = gather

    @staticmethod
    def fill_synthetic(n_rows_total: int, dim: int, rank: int, world: int, seed: int, device) -> torch.Tensor:
        """This rank's shard of the host-replayable synthetic table (synth.hash_table_rows gives the same values)."""
        from . import _lib
        n_local = (n_rows_total - rank + world - 1) // world
        out = torch.empty((n_local, dim), dtype=torch.float32, device=device)
        b = _lib.dl(out)
        _lib.check(_lib.load().rs_hash_table_fill(b.ptr, rank, world, seed, _lib.stream_ptr()))
        return out

    def enable_peer_exchange(self, max_batch: int, dedup: bool = True, native: bool = True) -> None:
        """Switch lookup() to the peer-memory path (csrc/peer.cu): both all-to-alls become stores into NVLink peer
        memory with device-known counts, the owner-side gather writes straight into the requester's region, hot ids
        are requested once per batch (dedup) -- no NCCL, no host synchronisation, capturable in a CUDA graph.
        Collective: every rank calls it with the same max_batch (regions: world * max_batch * (8 + 4 * dim) bytes)."""
        from . import _lib
        dev = self.local_table.device
        G, bm = self.world, int(max_batch)
        if native:
            # the whole lookup behind one C call (rs_embedding_a2a_lookup): regions and plan buffers live in the library
            import ctypes
            self._lib_mod, self._lib = _lib, _lib.load()
            self._comm = Comm.get(dev, self.group)
            h = ctypes.c_void_p()
            bt = _lib.dl(self.local_table)
            with torch.cuda.device(dev):
                _lib.check(self._lib.rs_sharded_table_create(self._comm._h, bt.ptr, bm, 1 if dedup else 0, ctypes.byref(h)))
            self._p2p = {"b_max": bm, "dedup": bool(dedup), "native": h}
            return
        self._p2p = {
            "b_max": bm, "dedup": bool(dedup),
            "ids": PeerExchange(16 + 8 * bm, dev, self.group), "rows": PeerExchange(bm * self.dim * 4, dev, self.group),
            "counts": torch.zeros((G,), dtype=torch.int64, device=dev),
            "stage": torch.zeros((G, bm), dtype=torch.int64, device=dev),
        }
        if dedup:
            H = 1 << max(4, (2 * bm - 1).bit_length())
            self._p2p.update(keys=torch.empty((H,), dtype=torch.int64, device=dev), vals=torch.empty((H,), dtype=torch.int32, device=dev),
                             slot_of=torch.empty((bm,), dtype=torch.int32, device=dev))
        self._p2p["rows_view"] = self._p2p["rows"].region.view(torch.float32).view(G * bm, self.dim)
        self._lib_mod, self._lib = _lib, _lib.load()

    def close(self) -> None:
        p = getattr(self, "_p2p", None)
        if p and "native" in p:
            torch.cuda.synchronize()
            with torch.cuda.device(self.local_table.device):
                self._lib_mod.check(self._lib.rs_sharded_table_free(p["native"]))
            self._p2p = None
            self._comm.release()
            self._comm = None
        elif p:
            torch.cuda.synchronize()
            p["rows_view"] = None
            p["ids"].close()
            p["rows"].close()
            self._p2p = None

    def last_request_counts(self) -> List[int]:
        """Rows this rank requested from each owner in its last peer-memory lookup (with de-duplication their sum is the
        number of distinct ids of the batch); synchronises the current stream."""
        import ctypes
        p = self._p2p
        if "native" in p:
            out = (ctypes.c_int64 * self.world)()
            self._lib_mod.check(self._lib.rs_sharded_table_last_counts(p["native"], out, self._lib_mod.stream_ptr()))
            return list(out)
        return p["counts"].cpu().tolist()

    def _lookup_p2p(self, ids: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        p, L, lib = self._p2p, self._lib_mod, self._lib
        B = int(ids.shape[0])
        if B > p["b_max"]:
            raise ValueError(f"batch {B} exceeds the max_batch {p['b_max']} the peer regions were sized for")
        st = L.stream_ptr()
        if out is None:
            out = torch.empty((B, self.dim), dtype=torch.float32, device=ids.device)
        if "native" in p:
            b_ids, b_out = L.dl(ids), L.dl(out)
            L.check(lib.rs_embedding_a2a_lookup(p["native"], b_ids.ptr, b_out.ptr, st))
            return out
        b_ids, b_counts, b_stage, b_out = L.dl(ids), L.dl(p["counts"]), L.dl(p["stage"]), L.dl(out)
        h_ids, h_rows = p["ids"]._h, p["rows"]._h
        if p["dedup"]:
            b_k, b_v, b_s = L.dl(p["keys"]), L.dl(p["vals"]), L.dl(p["slot_of"])
            L.check(lib.rs_a2a_plan_peer(b_ids.ptr, self.world, 1, b_k.ptr, b_v.ptr, b_s.ptr, b_counts.ptr, b_stage.ptr, None, st))
        else:
            inv = torch.empty((B,), dtype=torch.int64, device=ids.device)
            b_inv = L.dl(inv)
            L.check(lib.rs_a2a_plan_peer(b_ids.ptr, self.world, 0, None, None, None, b_counts.ptr, b_stage.ptr, b_inv.ptr, st))
        L.check(lib.rs_peer_a2a_ids(h_ids, b_stage.ptr, b_counts.ptr, st))
        b_table = L.dl(self.local_table)
        L.check(lib.rs_peer_a2a_rows(h_ids, h_rows, b_table.ptr, st))      # owner-side gather fused with the way back
        if p["dedup"]:
            L.check(lib.rs_peer_a2a_restore(h_rows, b_ids.ptr, b_s.ptr, b_v.ptr, None, b_out.ptr, st))
        else:
            L.check(lib.rs_peer_a2a_restore(h_rows, b_ids.ptr, None, None, b_inv.ptr, b_out.ptr, st))
        return out

    def lookup(self, ids: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """ids: int64 [B] global rows (this rank's part of the batch) -> float32 [B, dim], in batch order.
        out (peer-memory path only): a preallocated result tensor."""
        G = self.world
        if getattr(self, "_p2p", None):
            return self._lookup_p2p(ids, out)
        if G == 1:
            return self._gather(self.local_table, ids)
        counts = self._count(ids, G)
        recv_counts_t = torch.empty_like(counts)
        dist.all_to_all_single(recv_counts_t, counts, group=self.group)
        send_counts = counts.cpu().tolist()
        recv_counts = recv_counts_t.cpu().tolist()
        base = torch.tensor([sum(send_counts[:g]) for g in range(G)], dtype=torch.int64).to(ids.device)
        send, perm = self._scatter(ids, G, base)
        recv_rows = torch.empty((sum(recv_counts),), dtype=torch.int64, device=ids.device)
        dist.all_to_all_single(recv_rows, send, output_split_sizes=recv_counts, input_split_sizes=send_counts, group=self.group)
        found = self._gather(self.local_table, recv_rows)                    # [n_recv, dim] on the owner
        back = torch.empty((int(ids.shape[0]), self.dim), dtype=torch.float32, device=ids.device)
        dist.all_to_all_single(back, found, output_split_sizes=send_counts, input_split_sizes=recv_counts, group=self.group)
        return self._gather(back, perm)


_SIDE_STREAMS = {}


def lookup_many(pairs):
    """[(ShardedEmbedding, ids), ...] -> [rows, ...].  A peer-memory lookup is a chain of small kernels and two NVLink
    round trips -- bound by latency, not by bandwidth -- so the tables are looked up CONCURRENTLY, each on its own
    side stream (their regions are independent); the results are joined back onto the current stream.  Capturable
    in a CUDA graph (the side streams become parallel branches).  Tables on the NCCL path are looked up one after
    the other: collectives of one communicator must be issued in the same order on every rank."""
    if not all(getattr(emb, "_p2p", None) for emb, _ in pairs) or len(pairs) < 2:
        return [emb.lookup(ids) for emb, ids in pairs]
    dev = pairs[0][1].device
    cur = torch.cuda.current_stream(dev)
    streams = _SIDE_STREAMS.setdefault(dev.index, [])
    while len(streams) < len(pairs):
        streams.append(torch.cuda.Stream(device=dev))
    # results are allocated on the CURRENT stream (the side streams only fill them), so that freeing them later
    # needs no cross-stream bookkeeping in the caching allocator
    outs = [torch.empty((int(ids.shape[0]), emb.dim), dtype=torch.float32, device=dev) for emb, ids in pairs]
    for (emb, ids), out, st in zip(pairs, outs, streams):
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            emb.lookup(ids, out=out)
    for st in streams[:len(pairs)]:
        cur.wait_stream(st)
    return outs

