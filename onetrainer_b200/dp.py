"""Data parallelism for the hot path — net-new: the reference is single-process, single-device
(/root/reference/modules/trainer/BaseTrainer.py:33-34 builds exactly one train_device).

One process per GPU (torchrun / torch.distributed.run), NCCL over NVLink 5 / NVSwitch on GPUs,
gloo on CPU for the host-logic tests.  The path shards by batch with exactly ONE exchange step
(SURVEY.md §8e): rank k owns global samples [k*b, (k+1)*b); everything before the optimizer is
per-sample; gradients of the trainable tensors are summed across ranks and averaged (1/n is folded
into the grad-norm / optimizer kernels as `grad_prescale`, so no extra pass over the gradients).

RNG: the reference seeds a fresh generator with `global_step` on every rank
(BaseStableDiffusionXLSetup.py:189-191); naive replication would draw identical noise everywhere.
Here every draw is keyed by the GLOBAL sample index, so an N-GPU run at per-GPU batch b draws
bit-for-bit what a 1-GPU run at batch N*b draws.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def is_initialized() -> bool:
    return dist.is_available() and dist.is_initialized()


def rank() -> int:
    return dist.get_rank() if is_initialized() else 0


def world_size() -> int:
    return dist.get_world_size() if is_initialized() else 1


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """Join the process group described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT.
    Returns (rank, local_rank, world_size).  No-op for a single process."""
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rk = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    if ws > 1 and not is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(lr)
            dist.init_process_group(backend=backend, rank=rk, world_size=ws, device_id=torch.device("cuda", lr))
        else:
            dist.init_process_group(backend=backend, rank=rk, world_size=ws)
    return rk, lr, ws


def global_sample_base(local_batch_size: int) -> int:
    """Global index of this rank's first sample: rank * b."""
    return rank() * int(local_batch_size)


_replicated = {"seed": None, "calls": 0}


def replicated_seed() -> int:
    """Rank 0's `torch.initial_seed()` on every rank (broadcast once; collective on first use).  See replicated_key()."""
    if _replicated["seed"] is None:
        seed = torch.initial_seed() & (2 ** 63 - 1)
        if world_size() > 1:
            t = torch.tensor([seed], dtype=torch.int64, device="cuda" if dist.get_backend() == "nccl" else "cpu")
            dist.broadcast(t, src=0)
            seed = int(t.item())
        _replicated["seed"] = seed
    return _replicated["seed"]


def replicated_key() -> tuple[int, int]:
    """A Philox (seed, offset) that is THE SAME on every rank for the k-th call: rank 0's `torch.initial_seed()`, broadcast
    once (collective: every rank must make its first call), and a call counter.  For draws that must not differ between
    replicas whatever per-rank seeding the application uses — stochastic rounding of replicated bf16 parameters by the
    stand-alone helpers (modules/util/bf16_stochastic_rounding.py): replicas make the same calls in the same order, so
    they round identically and do not drift apart.  (Noise and timesteps are the opposite case: keyed per global sample.)"""
    seed = replicated_seed()
    _replicated["calls"] += 1
    return seed, 4 * _replicated["calls"]


class GradientBucket:
    """Flat view of the trainable tensors' gradients for the one exchange step.

    When the gradients already live in one arena (LoRAModuleWrapper's GradArena) the allreduce
    runs in place on it; otherwise they are gathered into a flat buffer with one
    torch._foreach_copy_ and scattered back.  The collective is a SUM; the 1/n average is applied
    by the consumer (`grad_prescale`) — no extra read/write of the gradients.
    """

    def __init__(self, arena: torch.Tensor | None = None):
        self.arena = arena
        self._flat = None
        self._exchange: GradExchange | None = None
        self.norm_partials: torch.Tensor | None = None      # squared-norm partials of the last exchange (K-r path only)

    def allreduce(self, grads: list[torch.Tensor]) -> float:
        """Sum gradients over ranks in place.  Returns the prescale (1/world_size) to apply."""
        n = world_size()
        self.norm_partials = None
        if n == 1:
            return 1.0
        if self.arena is not None and all(_is_view_of(g, self.arena) for g in grads):
            if is_symmetric(self.arena):
                # K-r: reduce-scatter + all-gather + squared norm in one kernel over NVLink peer memory (csrc/exchange.cu)
                if self._exchange is None:
                    self._exchange = GradExchange(self.arena)
                self.norm_partials = self._exchange.exchange(1.0 / n)
            else:
                dist.all_reduce(self.arena, op=dist.ReduceOp.SUM)
            return 1.0 / n
        total = sum(g.numel() for g in grads)
        if self._flat is None or self._flat.numel() != total or self._flat.dtype != grads[0].dtype:
            self._flat = torch.empty(total, dtype=grads[0].dtype, device=grads[0].device)
        views, off = [], 0
        for g in grads:
            views.append(self._flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        torch._foreach_copy_(views, grads)
        dist.all_reduce(self._flat, op=dist.ReduceOp.SUM)
        torch._foreach_copy_(grads, views)
        return 1.0 / n


# ---------------------------------------------------------------------------------------------------------------
# symmetric (peer-mapped + NVLS multicast) memory for the exchange kernel
# ---------------------------------------------------------------------------------------------------------------
_symmetric: dict[int, int] = {}        # data_ptr -> bytes of every live symmetric_zeros() allocation


def peer_exchange_wanted() -> bool:
    """OTB_PEER_EXCHANGE=0 keeps the LoRA gradient exchange on ncclAllReduce (A/B runs); default on for GPU ranks."""
    return world_size() > 1 and torch.cuda.is_available() and os.environ.get("OTB_PEER_EXCHANGE", "1") != "0"


def sharded_optimizer_wanted() -> bool:
    """Full fine-tuning under data parallelism: reduce-scatter -> sharded AdamW/EMA -> all-gather with this repo's own
    kernels (ShardedExchange) instead of the bucketed ncclAllReduce + replicated optimizer.  OTB_SHARDED_OPT=0 keeps the
    replicated form (A/B runs)."""
    return world_size() > 1 and torch.cuda.is_available() and os.environ.get("OTB_SHARDED_OPT", "1") != "0"


def symmetric_zeros(numel: int, dtype: torch.dtype, device) -> torch.Tensor:
    """A zero-filled flat buffer from torch's symmetric-memory allocator: the same allocation exists on every rank and,
    after the rendezvous, is peer-mapped (NVLink loads/stores) and mapped as one NVLS multicast object.  Every rank
    must allocate the same sizes in the same order.  Plumbing only: the kernels that use it are ours."""
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(int(numel), dtype=dtype, device=torch.device(device))
    t.zero_()
    _symmetric[t.data_ptr()] = t.numel() * t.element_size()
    return t


def is_symmetric(t: torch.Tensor) -> bool:
    return t.is_cuda and _symmetric.get(t.data_ptr(), 0) >= t.numel() * t.element_size()


class GradExchange:
    """Host side of K-r (include/otb200.h: otb_grad_exchange) for one flat fp32 gradient arena.

    Construction is collective (three rendezvous + one barrier) and happens once, outside any graph capture;
    `exchange()` is three launches on the current stream, capturable, and returns the squared-norm partials
    (float[world * blocks], identical on every rank) for otb_grad_norm_from_partials.
    mode: "p2p" (peer loads / stores; 2, 4 or 8 ranks) or "nvls" (multimem through the switch); default from
    OTB_KR_MODE, else the faster one measured for the world size (p2p on 2 ranks, nvls from 4)."""

    def __init__(self, arena: torch.Tensor, group=None, mode: str | None = None):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        from . import _native as N
        if arena.dtype != torch.float32 or arena.numel() % 4 != 0 or not is_symmetric(arena):
            raise ValueError("GradExchange: the arena must be a symmetric_zeros() fp32 buffer with a multiple of 4 elements")
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("GradExchange: the rendezvous cannot be captured; run one eager step first")
        group = group if group is not None else dist.group.WORLD
        rk, ws = dist.get_rank(group), dist.get_world_size(group)
        if not 2 <= ws <= N.MAX_RANKS:
            raise ValueError(f"GradExchange: world size {ws} outside 2..{N.MAX_RANKS}")
        # measured on B200 / NVSwitch, 340 MB arena (profiles/r01d_logs/nvls_check*.log): 2 ranks p2p 524 us, nvls 889 us,
        # ncclAllReduce 650 us; 4 ranks 806 / 780 / 855 us; 8 ranks 961 / 757 / 885 us
        mode = mode or os.environ.get("OTB_KR_MODE") or ("p2p" if ws == 2 else "nvls")
        if mode not in ("p2p", "nvls"):
            raise ValueError(f"GradExchange: unknown mode {mode!r}")
        dev = arena.device
        self.arena, self.mode = arena, mode
        self.blocks = int(N.lib().otb_grad_exchange_blocks(arena.numel(), ws))
        self.partials = symmetric_zeros(ws * self.blocks, torch.float32, dev)
        self.flags = symmetric_zeros(2 * N.MAX_RANKS, torch.int32, dev)
        tensors = (arena, self.partials, self.flags)
        handles = [symm.rendezvous(t, group) for t in tensors]

        def peers(t, h):
            arr = (C.c_void_p * N.MAX_RANKS)()
            for r in range(ws):
                arr[r] = h.buffer_ptrs[r] + (t.data_ptr() - h.buffer_ptrs[rk])
            return arr

        def multicast(t, h):
            if mode != "nvls":
                return None
            if not h.multicast_ptr:
                raise RuntimeError("GradExchange: this allocation has no NVLS multicast mapping (no NVSwitch multicast support "
                                   "on this node); use OTB_KR_MODE=p2p, or OTB_PEER_EXCHANGE=0 for ncclAllReduce")
            return h.multicast_ptr + (t.data_ptr() - h.buffer_ptrs[rk])

        self._handles = handles                     # keep the mappings alive
        self._args = N.GradExchangeArgs(
            peer_grads=peers(arena, handles[0]), peer_partials=peers(self.partials, handles[1]), peer_flags=peers(self.flags, handles[2]),
            mc_grads=multicast(arena, handles[0]), mc_partials=multicast(self.partials, handles[1]), n_elems=arena.numel(),
            rank=rk, world=ws, blocks=self.blocks, mode=N.EXCHANGE_NVLS if mode == "nvls" else N.EXCHANGE_P2P, grad_prescale=1.0)
        torch.cuda.synchronize(dev)
        dist.barrier(group)                         # every rank's flag words are zero before anyone raises one

    def exchange(self, grad_prescale: float) -> torch.Tensor:
        import ctypes as C
        from . import _native as N, ops
        self._args.grad_prescale = float(grad_prescale)
        N.check(N.lib().otb_grad_exchange(C.byref(self._args), ops._stream()), "otb_grad_exchange")
        ops._count(3)
        return self.partials


class FlatGradArena:
    """Persistent, flat gradient storage for a parameter list (full fine-tuning, config C4).

    `p.grad` of every parameter becomes a view of ONE buffer in the parameter dtype and stays that view for the whole
    run (autograd accumulates into an existing `.grad` in place).  That gives the data-parallel exchange the same
    zero-copy in-place allreduce the LoRA arena has, and it gives a captured CUDA graph gradient addresses that do
    not move between replays — with `zero_grad(set_to_none=True)` autograd allocates new gradient tensors every step.
    After the optimizer step call `zero()` (one memset) instead of `optimizer.zero_grad(set_to_none=True)`.
    """

    def __init__(self, parameters: list[torch.Tensor]):
        if not parameters:
            raise ValueError("FlatGradArena: empty parameter list")
        dt, dev = parameters[0].dtype, parameters[0].device
        if any(p.dtype != dt or p.device != dev for p in parameters):
            raise ValueError("FlatGradArena: parameters must share dtype and device")
        align = 16 // parameters[0].element_size()
        total, offsets = 0, []
        for p in parameters:
            offsets.append(total)
            total += (p.numel() + align - 1) // align * align           # every slice 16-byte aligned
        self.flat = torch.zeros((total,), dtype=dt, device=dev)
        self.parameters = list(parameters)
        for p, off in zip(parameters, offsets):
            p.grad = self.flat[off:off + p.numel()].view(p.shape)

    def zero(self) -> None:
        self.flat.zero_()

    def grads(self) -> list[torch.Tensor]:
        return [p.grad for p in self.parameters]


class OverlappedArenaReducer:
    """The one exchange step of a FlatGradArena, overlapped with backward (full fine-tuning, config C4).

    The 4.4 GB bf16 gradient arena of SDXL costs ~9 ms of allreduce on 2 GPUs and more on 8 when it runs after
    backward.  Gradients become final in (roughly) reverse parameter order, so the arena is cut into contiguous buckets;
    a post-accumulate-grad hook counts each bucket's parameters down and, when the last one lands, issues that bucket's
    in-place allreduce on a communication stream forked from the backward stream.  `finish()` joins the streams and
    returns the 1/n prescale.  Works the same inside a CUDA-graph capture (fork / join of captured streams).
    With gradient accumulation, `arm(last_micro_step=False)` keeps the hooks quiet until the last micro-batch.
    """

    def __init__(self, arena: FlatGradArena, n_buckets: int | None = None):
        if n_buckets is None:
            n_buckets = int(os.environ.get("OTB_REDUCE_BUCKETS", "12"))
        self.arena = arena
        params = arena.parameters
        total = arena.flat.numel()
        bounds, self._bucket_of, counts = [0], {}, []
        target = max(1, total // max(1, n_buckets))
        start_ptr = arena.flat.data_ptr()
        esz = arena.flat.element_size()
        cur = 0
        for i, p in enumerate(params):
            off = (p.grad.data_ptr() - start_ptr) // esz
            if off - bounds[-1] >= target and cur > 0:
                bounds.append(off)
                counts.append(cur)
                cur = 0
            self._bucket_of[id(p)] = len(bounds) - 1
            cur += 1
        bounds.append(total)
        counts.append(cur)
        self._slices = [arena.flat[bounds[b]:bounds[b + 1]] for b in range(len(counts))]
        self._counts = counts
        self._left = list(counts)
        self._armed = False
        self._comm = None
        self._issued = 0
        for p in params:
            p.register_post_accumulate_grad_hook(self._on_grad)

    def arm(self, last_micro_step: bool = True) -> None:
        self._left = list(self._counts)
        self._armed = last_micro_step and world_size() > 1
        self._issued = 0

    def _on_grad(self, p: torch.Tensor) -> None:
        if not self._armed:
            return
        b = self._bucket_of[id(p)]
        self._left[b] -= 1
        if self._left[b] == 0:
            self._reduce(b)

    def _reduce(self, b: int) -> None:
        if not self.arena.flat.is_cuda:                   # host tensors (gloo tests): no streams to fork
            dist.all_reduce(self._slices[b], op=dist.ReduceOp.SUM)
            self._issued += 1
            return
        main = torch.cuda.current_stream(self.arena.flat.device)
        if self._comm is None:
            self._comm = torch.cuda.Stream(device=self.arena.flat.device)
        self._comm.wait_stream(main)                      # the kernels that produced this bucket's gradients
        with torch.cuda.stream(self._comm):
            dist.all_reduce(self._slices[b], op=dist.ReduceOp.SUM)
        self._issued += 1

    def finish(self) -> float:
        """Reduce whatever has not been reduced (parameters without a gradient this step), join, return the prescale."""
        n = world_size()
        if n == 1:
            return 1.0
        if not self._armed:
            return 1.0
        for b, left in enumerate(self._left):
            if left > 0:
                self._left[b] = 0
                self._reduce(b)
        if self._comm is not None and self.arena.flat.is_cuda:
            torch.cuda.current_stream(self.arena.flat.device).wait_stream(self._comm)
        self._armed = False
        return 1.0 / n


class ShardedExchange:
    """K-z, host side: reduce-scatter -> sharded optimizer -> all-gather for full fine-tuning (config C4; SURVEY.md §8f
    rank 3), over NVLink peer memory with this repo's own kernels (include/otb200.h: otb_shard_reduce / otb_shard_gather).

    Construction (collective, outside any graph capture) re-homes the parameters AND their gradients into two flat
    symmetric-memory arenas (same offsets in both), cuts them into contiguous tensor-aligned buckets, and gives bucket b
    to rank b % world.  Per step:

        arm()                 before backward
        [hooks]               overlap=True only (off by default — measured slower, see below): when a bucket's last gradient lands,
                              non-owners signal the owner, the owner waits for every peer and reduces the bucket on a
                              communication stream while backward goes on
        finish() -> 1/n       joins; after it every rank holds the complete table of squared-norm partials
                              (`norm_partials`) and the owner's replica holds the summed gradients of its buckets
        optimizer.step()      an optimizer built over `owned_parameters()` only: 1/n of the state, EMA and HBM traffic
        gather()              the owners push their updated parameter ranges into every replica; barrier

    Same interface as OverlappedArenaReducer (arm / finish), so the trainer treats both alike.  The stochastic-rounding
    stream stays keyed by the GLOBAL element offset (`rng_base_of`), so the sharded update is bit-identical to the
    replicated one given the same summed gradients."""

    def __init__(self, parameters: list[torch.Tensor], n_buckets: int | None = None, group=None, mode: str | None = None,
                 overlap: bool | None = None):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        from . import _native as N
        if not parameters:
            raise ValueError("ShardedExchange: empty parameter list")
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("ShardedExchange: the rendezvous cannot be captured")
        dt, dev = parameters[0].dtype, parameters[0].device
        if any(p.dtype != dt or p.device != dev for p in parameters) or dt not in (torch.bfloat16, torch.float32):
            raise ValueError("ShardedExchange: parameters must share one device and one dtype (bf16 or fp32)")
        group = group if group is not None else dist.group.WORLD
        rk, ws = dist.get_rank(group), dist.get_world_size(group)
        if not 2 <= ws <= N.MAX_RANKS:
            raise ValueError(f"ShardedExchange: world size {ws} outside 2..{N.MAX_RANKS}")
        mode = mode or os.environ.get("OTB_KZ_MODE") or ("p2p" if ws == 2 else "nvls")
        if mode == "p2p" and ws not in (2, 4, 8):
            mode = "nvls"
        self.mode, self.rank, self.world = mode, rk, ws
        # Per-bucket reduce DURING backward (hooks) or all buckets after it?  Measured on C4, 142.0 ms/step on one GPU:
        #   2 B200s, P2P  (profiles/r02e_logs): after backward 145.5 ms; during backward 149.6 ms (296 blocks), 155 / 172 / 235 ms
        #                 with 32 / 16 / 8 blocks — peer loads need many SMs to keep up, and backward needs the same SMs;
        #   8 B200s, NVLS (profiles/r02p_logs): after backward 151.8 ms; during backward 148.6 ms — the switch does the adds, the
        #                 reduce is one request stream and disturbs backward less than it saves.
        #   4 B200s, NVLS (profiles/r02x_logs, r02y_logs): after backward 148.3 ms, during backward 150.0 ms (P2P after: 148.8 ms).
        # Default: overlap from 8 ranks up in NVLS mode, otherwise reduce after backward; OTB_KZ_OVERLAP=0|1 overrides.
        env = os.environ.get("OTB_KZ_OVERLAP")
        self.overlap = ((env != "0") if env is not None else (mode == "nvls" and ws >= 8)) if overlap is None else overlap
        self.parameters = list(parameters)
        offsets, total, bounds, first_param = self.plan([p.numel() for p in parameters], ws, n_buckets)
        self.param_flat = symmetric_zeros(total, dt, dev)
        self.grad_flat = symmetric_zeros(total, dt, dev)
        with torch.no_grad():
            for p, off in zip(parameters, offsets):
                view = self.param_flat[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                p.grad = self.grad_flat[off:off + p.numel()].view(p.shape)
        self.flat = self.grad_flat                      # FlatGradArena's interface (zero(), grads())
        self.offsets = offsets
        self.rng_base_of = {id(p): off for p, off in zip(parameters, offsets)}     # 8-aligned (align >= 4; bf16: 8)
        self.bounds = bounds
        self.n_buckets = len(bounds) - 1
        self.owner = [b % ws for b in range(self.n_buckets)]
        self._bucket_of = {}
        self._counts = []
        for b in range(self.n_buckets):
            for i in range(first_param[b], first_param[b + 1]):
                self._bucket_of[id(parameters[i])] = b
            self._counts.append(first_param[b + 1] - first_param[b])
        self._owned = [p for p in parameters if self.owner[self._bucket_of[id(p)]] == rk]
        self.blocks = min(4096, max(1, int(os.environ.get("OTB_KZ_BLOCKS", str(4 * torch.cuda.get_device_properties(dev).multi_processor_count)))))
        self.partials = symmetric_zeros(self.n_buckets * self.blocks, torch.float32, dev)
        self.flags = symmetric_zeros(N.PEER_PHASES * N.MAX_RANKS, torch.int32, dev)              # barriers
        self.ready = symmetric_zeros(self.n_buckets * N.MAX_RANKS, torch.int32, dev)             # per-bucket "produced" words
        tensors = (self.grad_flat, self.param_flat, self.partials, self.flags, self.ready)
        handles = [symm.rendezvous(t, group) for t in tensors]
        self._handles = handles

        def peers(t, h):
            arr = (C.c_void_p * N.MAX_RANKS)()
            for r in range(ws):
                arr[r] = h.buffer_ptrs[r] + (t.data_ptr() - h.buffer_ptrs[rk])
            return arr

        def multicast(t, h):
            if mode != "nvls":
                return None
            if not h.multicast_ptr:
                raise RuntimeError("ShardedExchange: no NVLS multicast mapping on this node; use OTB_KZ_MODE=p2p")
            return h.multicast_ptr + (t.data_ptr() - h.buffer_ptrs[rk])

        common = dict(dtype=N.dtype_code(dt), rank=rk, world=ws, mode=N.EXCHANGE_NVLS if mode == "nvls" else N.EXCHANGE_P2P,
                      blocks=self.blocks, grad_prescale=1.0 / ws)
        self._reduce_args = N.ShardExchangeArgs(peer_base=peers(self.grad_flat, handles[0]), peer_partials=peers(self.partials, handles[2]),
                                                mc_base=multicast(self.grad_flat, handles[0]), mc_partials=multicast(self.partials, handles[2]),
                                                **common)
        self._gather_args = N.ShardExchangeArgs(peer_base=peers(self.param_flat, handles[1]), peer_partials=(C.c_void_p * N.MAX_RANKS)(),
                                                mc_base=multicast(self.param_flat, handles[1]), mc_partials=None, **common)
        # K-c with the all-gather folded in (otb_adamw_sr_ema_multi_peers): byte offsets from my replica of the parameter arena
        ps = N.PeerStore(rank=rk, world=ws, mode=N.EXCHANGE_NVLS if mode == "nvls" else N.EXCHANGE_P2P)
        mine = handles[1].buffer_ptrs[rk]
        for r in range(ws):
            ps.delta[r] = handles[1].buffer_ptrs[r] - mine
        ps.mc_delta = (handles[1].multicast_ptr - mine) if mode == "nvls" else 0
        self.peer_store = ps if os.environ.get("OTB_KZ_FUSED_GATHER", "1") != "0" else None
        self.gather_is_fused = False        # set by the owner of the optimizer once its driver carries `peer_store`
        self._flag_ptrs = peers(self.flags, handles[3])
        self._ready_ptrs = peers(self.ready, handles[4])
        self._my_ready = self.ready.data_ptr()
        self.norm_partials = self.partials
        self._left = list(self._counts)
        self._armed = False
        self._comm = None
        self._reduced = [False] * self.n_buckets
        for p in parameters:
            p.register_post_accumulate_grad_hook(self._on_grad)
        torch.cuda.synchronize(dev)
        dist.barrier(group)

    @staticmethod
    def plan(numels: list[int], world: int, n_buckets: int | None = None):
        """Arena layout and bucket cut for a parameter list (pure host logic).  Every tensor starts on an 8-element
        boundary (16-byte vectors for bf16; the 8-element groups of the SR Philox stream, so offsets equal the replicated
        optimizer table's `rng_base`).  Buckets are contiguous, tensor-aligned, about total / n_buckets elements each, a
        multiple of the world size of them where the tensor count allows; bucket b belongs to rank b % world.
        Returns (offsets, total, bounds, first_param): bucket b covers elements [bounds[b], bounds[b+1]) = parameters
        [first_param[b], first_param[b+1])."""
        align = 8
        offsets, total = [], 0
        for n in numels:
            offsets.append(total)
            total += (n + align - 1) // align * align
        if n_buckets is None:
            n_buckets = int(os.environ.get("OTB_KZ_BUCKETS", str(3 * world if world > 2 else 8)))
        n_buckets = max(world, (n_buckets + world - 1) // world * world)
        if len(numels) >= world:
            n_buckets = min(n_buckets, max(world, len(numels) // world * world))
        target = max(1, total // n_buckets)
        bounds, first_param = [0], [0]
        for i, off in enumerate(offsets):
            if off - bounds[-1] >= target and len(bounds) < n_buckets and i > first_param[-1]:
                bounds.append(off)
                first_param.append(i)
        bounds.append(total)
        first_param.append(len(numels))
        return offsets, total, bounds, first_param

    # FlatGradArena's interface -----------------------------------------------------------------------------------------
    def zero(self) -> None:
        self.grad_flat.zero_()

    def grads(self) -> list[torch.Tensor]:
        return [p.grad for p in self.parameters]

    def owned_parameters(self) -> list[torch.Tensor]:
        return list(self._owned)

    def owned_elements(self) -> int:
        return sum(self.bounds[b + 1] - self.bounds[b] for b in range(self.n_buckets) if self.owner[b] == self.rank)

    # the step ----------------------------------------------------------------------------------------------------------
    def arm(self, last_micro_step: bool = True) -> None:
        self._left = list(self._counts)
        self._reduced = [False] * self.n_buckets
        self._armed = last_micro_step

    def _on_grad(self, p: torch.Tensor) -> None:
        if not self._armed or not self.overlap:
            return
        b = self._bucket_of[id(p)]
        self._left[b] -= 1
        if self._left[b] == 0:
            main = torch.cuda.current_stream(self.grad_flat.device)
            if self._comm is None:
                self._comm = torch.cuda.Stream(device=self.grad_flat.device)
            self._comm.wait_stream(main)                  # the kernels that produced this bucket's gradients
            with torch.cuda.stream(self._comm):
                self._handshake_and_reduce(b)

    def _handshake_and_reduce(self, b: int) -> None:
        """Per-bucket hand-shake (non-owners raise their word in the owner's table; the owner consumes them) and, on the
        owner, the reduce of the bucket.  Every rank runs this for every bucket in the same order."""
        from . import _native as N, ops
        lib = N.lib()
        own = self.owner[b]
        word = b * N.MAX_RANKS
        if own == self.rank:
            N.check(lib.otb_peer_wait(self._my_ready + 4 * word, self.rank, self.world, ops._stream()), "otb_peer_wait")
            self._launch_reduce(b)
            ops._count(2)
        else:
            N.check(lib.otb_peer_signal(self._ready_ptrs[own] + 4 * (word + self.rank), ops._stream()), "otb_peer_signal")
            ops._count(1)
        self._reduced[b] = True

    def _launch_reduce(self, b: int) -> None:
        import ctypes as C
        from . import _native as N, ops
        a = self._reduce_args
        a.begin, a.end, a.partial_slot = self.bounds[b], self.bounds[b + 1], b * self.blocks
        N.check(N.lib().otb_shard_reduce(C.byref(a), ops._stream()), "otb_shard_reduce")

    def _barrier(self, phase: int) -> None:
        from . import _native as N, ops
        N.check(N.lib().otb_peer_barrier(self._flag_ptrs, self.rank, self.world, phase, ops._stream()), "otb_peer_barrier")
        ops._count(1)

    def finish(self) -> float:
        """Reduce what the hooks have not reduced, join the communication stream, make every rank's partial table
        complete.  Returns the prescale (1/world) for the norm / optimizer."""
        if not self._armed:
            return 1.0
        dev = self.grad_flat.device
        main = torch.cuda.current_stream(dev)
        if self._comm is not None and self.overlap:
            main.wait_stream(self._comm)
        pending = [b for b in range(self.n_buckets) if not self._reduced[b]]
        if pending:
            if len(pending) == self.n_buckets:
                self._barrier(0)                          # nothing overlapped: ONE barrier says every backward is complete
                for b in pending:
                    if self.owner[b] == self.rank:
                        self._launch_reduce(b)
                        from . import ops
                        ops._count(1)
                    self._reduced[b] = True
            else:
                for b in pending:                         # parameters without a gradient this step left a bucket open
                    self._handshake_and_reduce(b)
        self._barrier(1)                                  # every owner's partials have landed in every table
        self._armed = False
        return 1.0 / self.world

    def measure(self, iters: int = 5) -> dict:
        """CUDA-event time of the two halves, NOT overlapped with anything (collective: every rank must call it; the
        gradients are summed in place again and again, so call it only when their values no longer matter): max over ranks."""
        dev = self.grad_flat.device
        keep = self.overlap
        self.overlap = False

        def timed(fn):
            fn()
            torch.cuda.synchronize(dev)
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize(dev)
            t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()

        def reduce():
            self.arm()
            self.finish()
        fused = self.gather_is_fused
        self.gather_is_fused = False                      # time the stand-alone gather kernel
        try:
            red, gat = timed(reduce), timed(self.gather)
        finally:
            self.overlap = keep
            self.gather_is_fused = fused
        esz = self.grad_flat.element_size()
        nbytes = self.grad_flat.numel() * esz
        moved = (self.world - 1) / self.world * nbytes
        return {"mode": self.mode, "buckets": self.n_buckets, "arena_gb": nbytes / 1e9, "reduce_ms": red, "gather_ms": gat,
                "reduce_gbs_in_per_gpu": moved / red / 1e6, "gather_gbs_out_per_gpu": moved / gat / 1e6,
                "gather_folded_into_optimizer": fused,
                "note": "reduce-scatter + squared norm / all-gather of the updated parameters, each timed alone with its barriers; in the "
                        "step " + ("the reduce of a bucket overlaps the rest of backward" if keep else "the reduce runs after backward")
                        + ("; the gather is folded into the optimizer launch (its parameter stores go to every replica)" if fused
                           else "; the gather runs after the optimizer")}

    def gather(self) -> None:
        """After the optimizer step on the owned tensors: push my parameter ranges into every replica."""
        import ctypes as C
        from . import _native as N, ops
        a = self._gather_args
        if not self.gather_is_fused:                      # else: the optimizer launch already stored into every replica
            for b in range(self.n_buckets):
                if self.owner[b] == self.rank:
                    a.begin, a.end, a.partial_slot = self.bounds[b], self.bounds[b + 1], 0
                    N.check(N.lib().otb_shard_gather(C.byref(a), ops._stream()), "otb_shard_gather")
                    ops._count(1)
        self._barrier(2)                                  # every replica is complete before the next forward reads it
        try:
            torch.autograd.graph.increment_version(self.parameters)
        except TypeError:
            for p in self.parameters:
                torch.autograd.graph.increment_version(p)


def _is_view_of(t: torch.Tensor, base: torch.Tensor) -> bool:
    lo = base.data_ptr()
    hi = lo + base.numel() * base.element_size()
    return lo <= t.data_ptr() and t.data_ptr() + t.numel() * t.element_size() <= hi


def average_scalar(x: torch.Tensor) -> torch.Tensor:
    """Mean of a (device) scalar over ranks — used only for the logged loss."""
    if world_size() == 1:
        return x
    y = x.clone()
    dist.all_reduce(y, op=dist.ReduceOp.SUM)
    return y / world_size()
