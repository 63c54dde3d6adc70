"""Data-parallel plumbing of the hot path.

The reference's only collective is Lightning DDP's gradient all-reduce (`strategy: ddp_find_unused_parameters_true`,
diffusionGS/configs/diffusionGS_rel.yaml:80): 460,391,424 fp32 gradients per step, default 25 MB buckets, then a local
grad-norm clip at 0.5 (diffusionGS_rel.yaml:77).  Here the gradients live in ONE flat, pre-allocated arena (each
`param.grad` is a view into it, so backward kernels write their results in place and nothing is copied or bucketed
at step time); the all-reduce is issued per DiT block in REVERSE order (the order the backward produces them) on a
side stream over NCCL / NVLink, and the clip is a single fused norm over the arena.  One process per GPU.

Everything is torch.distributed: `nccl` on GPUs, `gloo` in the CPU tests (world_size 2).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None, device=None):
    """Rendezvous from torchrun's RANK / WORLD_SIZE / MASTER_* (127.0.0.1 when unset)."""
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if not dist.is_initialized():
        kw = {}
        if backend == "nccl" and device is not None:
            kw["device_id"] = device
        dist.init_process_group(backend, **kw)
    return dist.get_rank(), dist.get_world_size()


def gradient_group(max_ctas=None):
    """Optional communicator for the gradient all-reduce alone, with NCCL's CTA count capped (DGS_NCCL_MAX_CTAS > 0).  The
    collective runs UNDER the backward (overlapped per block), so its kernels compete with the persistent tcgen05 GEMMs for
    SMs (weight-gradient GEMMs +1.5 ms per step at 2 GPUs).  Measured at 2 x B200, batch 4 per GPU (r2, gpurun call 14):
    cap 2 / 4 / 8 / 16 / none = 107.7 / 106.2 / 102.4 / 102.6 / 101.3 ms per step -- a slower exchange overlaps MORE kernels
    and costs more than it frees, so the default is NO cap (returns None = the default group)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1 or dist.get_backend() != "nccl":
        return None
    if max_ctas is None:
        max_ctas = int(os.environ.get("DGS_NCCL_MAX_CTAS", "0"))
    if max_ctas <= 0:
        return None
    opts = dist.ProcessGroupNCCL.Options()
    opts.config.max_ctas = max_ctas
    opts.config.min_ctas = 1
    return dist.new_group(backend="nccl", pg_options=opts)


def max_over_ranks(value: float, device="cpu") -> float:
    """Device-time aggregation rule of bench.py: a multi-GPU time is the MAX over ranks."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def whole_job_throughput(units_per_rank: float, seconds_local: float, device="cpu") -> float:
    """value = units processed by ALL ranks / max-over-ranks time (weak scaling, no data-path collective)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    return world * units_per_rank / max_over_ranks(seconds_local, device)


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced partition of a batch of independent samples over ranks (the path shards by sample)."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


class GradArena:
    """Flat gradient arena + bucketed (per transformer block, reverse order) all-reduce + fused clip."""

    def __init__(self, module: torch.nn.Module, dtype=torch.float32, bucket_key=None):
        params = [p for p in module.parameters() if p.requires_grad]
        names = {id(p): n for n, p in module.named_parameters()}
        self.params = params
        total = sum(p.numel() for p in params)
        dev = params[0].device
        self.flat = torch.zeros(total, dtype=dtype, device=dev)
        # bucket = "transformer.<i>" for block parameters, "other" for the rest (tokenizer, heads, embedder)
        key = bucket_key or (lambda n: ".".join(n.split(".")[:2]) if n.startswith("transformer.") else "other")
        self.buckets = {}  # name -> [start, end) in arena order
        off, last_base, last_name = 0, None, None
        for p in params:
            base = key(names[id(p)])
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            if base == last_base:  # contiguous run of the same key: extend
                self.buckets[last_name][1] = off + n
            else:                  # new run (a key that re-appears later gets a suffixed bucket of its own)
                last_name = base if base not in self.buckets else f"{base}#{off}"
                self.buckets[last_name] = [off, off + n]
                last_base = base
            off += n
        self.total = total
        self._stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None

    def zero_(self):
        self.flat.zero_()

    def reverse_bucket_order(self):
        """transformer.23 ... transformer.0, then the remaining parameters (what finishes last in the backward)."""
        blk = sorted((k for k in self.buckets if k.startswith("transformer.")),
                     key=lambda k: -int(k.split(".")[1].split("#")[0]))
        return blk + [k for k in self.buckets if not k.startswith("transformer.")]

    def bucket_block(self, name):
        """transformer block index of a bucket name, None for the non-block buckets."""
        return int(name.split(".")[1].split("#")[0]) if name.startswith("transformer.") else None

    def allreduce_issue_(self, group=None, gate=None, sync_main=True):
        """Enqueue the SUM all-reduce of every bucket in reverse order on the side stream and return at once.
        `gate(block_index_or_None)` is called (with the side stream current) right before a bucket's collective is
        enqueued: the overlapped training step passes a function that makes the side stream wait on the event the
        backward records when that block's gradients are final (dgs_dit_backward_ex / dgs_stream_wait_event), so bucket l
        travels over NVLink while blocks l-1 .. 0 are still being differentiated -- what torch DDP's autograd hooks do for
        the reference (diffusionGS_rel.yaml:80).  sync_main: the side stream first waits for everything already enqueued on
        the current stream (the un-overlapped form: the whole backward)."""
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._works, self._world = [], world
        if world == 1:
            return self
        if self._stream is not None and sync_main:
            self._stream.wait_stream(torch.cuda.current_stream(self.flat.device))
        ctx = torch.cuda.stream(self._stream) if self._stream is not None else _null()
        bf16 = self.transport_bf16()
        self._pending_back = []
        with ctx:
            prev = None
            for gate_block, a, b in self.reduce_plan():
                if gate is not None:
                    gate(gate_block)
                if bf16:  # optional bf16 transport: half the bytes on the wire, the sum of `world` bf16 values on arrival
                    self._half[a:b].copy_(self.flat[a:b])
                    w = dist.all_reduce(self._half[a:b], op=dist.ReduceOp.SUM, group=group, async_op=True)
                    if prev is not None:  # copy the PREVIOUS bucket back to fp32 on the side stream, under the backward
                        prev[0].wait()
                        self.flat[prev[1]:prev[2]].copy_(self._half[prev[1]:prev[2]])
                        self._pending_back.remove(prev)
                    prev = (w, a, b)
                    self._pending_back.append(prev)
                else:
                    w = dist.all_reduce(self.flat[a:b], op=dist.ReduceOp.SUM, group=group, async_op=True)
                self._works.append(w)
        return self

    def transport_bf16(self):
        """DGS_GRAD_TRANSPORT=bf16: gradients cross NVLink as bf16 (0.92 instead of 1.84 GB per step).  NOT the reference's
        numerics (DDP all-reduces fp32), hence off by default; the arena, the clip norm and AdamW stay fp32."""
        on = os.environ.get("DGS_GRAD_TRANSPORT", "fp32").lower() == "bf16" and self.flat.dtype == torch.float32
        if on and getattr(self, "_half", None) is None:
            self._half = torch.empty(self.total, dtype=torch.bfloat16, device=self.flat.device)
        return on

    def reduce_plan(self, blocks_per_call=None):
        """[(gate block, start, end)] in issue order: `blocks_per_call` consecutive transformer blocks (contiguous in the
        arena) travel as ONE collective, gated on the LAST of them to be differentiated (the lowest index); the non-block
        buckets follow, gated on the end of the backward (None).  One block per call = 24 collectives of 75 MB, each paying
        the latency of an 8-rank collective; more blocks per call = fewer, larger collectives and a longer un-overlapped
        tail.  Measured at 8 x B200, batch 4 per GPU (r2, profiles/r2_allreduce_sweep_n8.txt): 1 / 3 / 6 blocks per call =
        107.8 / 101.9 / 102.4 ms per step (exposed 7.3 / 3.9 / 4.7 ms); default 3 (DGS_AR_BLOCKS_PER_CALL)."""
        k = blocks_per_call or int(os.environ.get("DGS_AR_BLOCKS_PER_CALL", "3"))
        order = self.reverse_bucket_order()
        blk = [n for n in order if n.startswith("transformer.")]
        plan = []
        for i in range(0, len(blk), max(k, 1)):
            grp = blk[i:i + max(k, 1)]
            lo = min(self.buckets[n][0] for n in grp)
            hi = max(self.buckets[n][1] for n in grp)
            assert hi - lo == sum(self.buckets[n][1] - self.buckets[n][0] for n in grp), "blocks of a group must be contiguous"
            plan.append((self.bucket_block(grp[-1]), lo, hi))
        plan += [(None, *self.buckets[n]) for n in order if not n.startswith("transformer.")]
        return plan

    def allreduce_wait_(self, scale=True):
        """Make the current stream wait for the collectives of allreduce_issue_; scale=True divides by the world size
        here, scale=False leaves the SUM in the arena and returns 1/world for the consumer to fold in (the fused AdamW
        takes it as its gradient scale, saving a pass over 1.8 GB)."""
        works, world = getattr(self, "_works", []), getattr(self, "_world", 1)
        for w in works:
            w.wait()
        self._works = []
        for (_w, a, b) in getattr(self, "_pending_back", []):  # bf16 transport: the buckets not yet copied back to fp32
            self.flat[a:b].copy_(self._half[a:b])
        self._pending_back = []
        if world > 1 and scale:
            self.flat.mul_(1.0 / world)
            return 1.0
        return 1.0 / world if world > 1 else 1.0

    def allreduce_mean_(self, group=None):
        """sum over ranks / world, bucket by bucket in reverse order; async on a side stream on GPUs."""
        self.allreduce_issue_(group)
        self.allreduce_wait_(scale=True)
        return self

    def clip_grad_norm_(self, max_norm: float, eps: float = 1e-6) -> torch.Tensor:
        """torch.nn.utils.clip_grad_norm_ semantics (L2, scale = max_norm / (norm + eps), clamped to 1) on the arena.
        Identical on every rank after the all-reduce, so no further collective."""
        norm = torch.linalg.vector_norm(self.flat.float())
        self.flat.mul_(torch.clamp(max_norm / (norm + eps), max=1.0).to(self.flat.dtype))
        return norm


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
