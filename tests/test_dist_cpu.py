"""world_size-2 gloo tests (CPU) of the N>1 host logic: sample sharding, max-over-ranks timing, whole-job throughput
and the flat gradient arena's bucketed all-reduce + clip (row a17) against torch DDP-style averaging."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dgs_b200 import dist as dd


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dd.init_from_env("gloo")
    try:
        # --- timing / throughput aggregation
        t = dd.max_over_ranks(1.0 + rank)
        thr = dd.whole_job_throughput(10.0, 1.0 + rank)
        # --- flat arena all-reduce vs per-parameter reference
        torch.manual_seed(0)
        from dgs_b200.denoiser import DGSDenoiser
        model = DGSDenoiser(dict(patch_size=8, num_layers=3, width=1024))
        arena = dd.GradArena(model)
        g = torch.Generator().manual_seed(100 + rank)
        ref = []
        for p in model.parameters():
            v = torch.randn(p.shape, generator=g)
            p.grad.copy_(v)          # backward kernels would write here in place
            ref.append(v)
        assert all(p.grad.data_ptr() >= arena.flat.data_ptr() for p in model.parameters())
        arena.allreduce_mean_()
        # reference: average of the two ranks' tensors, computed independently
        other = []
        g2 = torch.Generator().manual_seed(100 + (1 - rank))
        for p in model.parameters():
            other.append(torch.randn(p.shape, generator=g2))
        ok = all(torch.allclose(p.grad, (a + b) / 2, atol=1e-6) for p, a, b in zip(model.parameters(), ref, other))
        expect_norm = torch.sqrt(sum((((a + b) / 2) ** 2).sum() for a, b in zip(ref, other)))
        norm = arena.clip_grad_norm_(0.5)
        after = torch.linalg.vector_norm(arena.flat)
        order = arena.reverse_bucket_order()
        ret[rank] = dict(t=t, thr=thr, ok=ok, norm=float(norm), expect=float(expect_norm), after=float(after),
                         order=order[:4], nb=len(order), total=arena.total,
                         shard=dd.shard_range(7, rank, world))
    finally:
        dist.destroy_process_group()


def test_gloo_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29611, ret), nprocs=world, join=True)
    r0, r1 = ret[0], ret[1]
    assert r0["t"] == r1["t"] == 2.0                      # max over ranks
    assert abs(r0["thr"] - 2 * 10.0 / 2.0) < 1e-12        # whole-job units / slowest rank
    assert r0["ok"] and r1["ok"]
    assert abs(r0["norm"] - r0["expect"]) < 1e-2 * r0["expect"] and abs(r0["norm"] - r1["norm"]) < 1e-3
    assert abs(r0["after"] - 0.5) < 1e-3                  # clipped to max_norm
    assert r0["order"][:3] == ["transformer.2", "transformer.1", "transformer.0"] and r0["nb"] == 5  # 3 blocks + the parameters before and after the stack
    assert r0["shard"] == (0, 4) and r1["shard"] == (4, 7)


def test_arena_layout_single_process():
    from dgs_b200.denoiser import DGSDenoiser
    m = DGSDenoiser(dict(patch_size=8, num_layers=2))
    a = dd.GradArena(m)
    assert a.total == sum(p.numel() for p in m.parameters())
    sizes = {k: b - s for k, (s, b) in a.buckets.items()}
    assert sizes["transformer.0"] == sizes["transformer.1"] == 18_889_728   # SURVEY 8e bucket size
    assert sum(sizes.values()) == a.total
    a.flat.fill_(1.0)
    assert all(float(p.grad.min()) == 1.0 for p in m.parameters())
    n = a.clip_grad_norm_(0.5)
    assert abs(float(n) - a.total ** 0.5) < 1e-3 * a.total ** 0.5


def _overlap_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dd.init_from_env("gloo")
    try:
        from dgs_b200.denoiser import DGSDenoiser
        torch.manual_seed(0)
        model = DGSDenoiser(dict(patch_size=8, num_layers=3))
        arena = dd.GradArena(model)
        arena.flat.fill_(float(rank + 1))
        gated = []
        os.environ["DGS_AR_BLOCKS_PER_CALL"] = "1"
        arena.allreduce_issue_(gate=gated.append, sync_main=False)   # the overlapped form: one gate per bucket, in issue order
        inv = arena.allreduce_wait_(scale=False)                     # SUM stays in the arena, 1/world goes to the consumer
        ret[rank] = dict(gated=gated, inv=inv, lo=float(arena.flat.min()), hi=float(arena.flat.max()))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_gated_issue_and_deferred_scale():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_overlap_worker, args=(world, 29613, ret), nprocs=world, join=True)
    for r in (ret[0], ret[1]):
        assert r["gated"] == [2, 1, 0, None, None]   # blocks in reverse order, then the two non-block buckets
        assert r["inv"] == 0.5 and r["lo"] == r["hi"] == 3.0


def _bf16_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      DGS_GRAD_TRANSPORT="bf16", DGS_AR_BLOCKS_PER_CALL="2")
    dd.init_from_env("gloo")
    try:
        from dgs_b200.denoiser import DGSDenoiser
        torch.manual_seed(0)
        model = DGSDenoiser(dict(patch_size=8, num_layers=3))
        arena = dd.GradArena(model)
        g = torch.Generator().manual_seed(7 + rank)
        mine = torch.randn(arena.total, generator=g)
        arena.flat.copy_(mine)
        plan = arena.reduce_plan()
        arena.allreduce_mean_()
        other = torch.randn(arena.total, generator=torch.Generator().manual_seed(7 + (1 - rank)))
        expect = (mine.bfloat16() + other.bfloat16()).float() / 2      # each rank's values rounded to bf16, summed in bf16
        err = float((arena.flat - (mine + other) / 2).norm() / ((mine + other) / 2).norm())
        ret[rank] = dict(plan=[(b, e - s) for b, s, e in plan], err=err,
                         close=bool(torch.allclose(arena.flat, expect, rtol=2e-2, atol=1e-3)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_bf16_transport_and_grouped_buckets():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_bf16_worker, args=(world, 29617, ret), nprocs=world, join=True)
    for r in (ret[0], ret[1]):
        # 3 blocks, 2 per call: {2, 1} gated on block 1, {0} gated on block 0, then the two non-block buckets
        assert [p[0] for p in r["plan"]] == [1, 0, None, None]
        assert r["plan"][0][1] == 2 * 18_889_728 and r["plan"][1][1] == 18_889_728
        assert r["close"] and 1e-4 < r["err"] < 1e-2      # bf16 rounding is visible but small
