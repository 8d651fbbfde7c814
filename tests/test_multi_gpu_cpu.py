"""N>1 host logic on CPU: world_size-2 gloo. The oracle stands in for the device kernels; the sharding/reduction logic is
the product's (python/vpt_shard.py), the same code bench.py and tools/mgpu_check.py run with libvpt + NCCL."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "real-time-path-tracing-voxel-blocks_b200", "python"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import common
    import oracle as O
    import vpt_scenes as S
    import vpt_shard
    O.set_threads(2)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    W, H, spp = 96, 64, 4
    inp = common.scene_inputs((2, 1, 2), noise_fn=O.perlin_noise_chunks, alias_fn=O.build_alias_table)
    o = common.setup(O.Oracle(W, H), inp, spp=spp, total=3, diffuse=1)
    cam = O.camera_from_scene(W, H, S.SCENE_CAMERA["position"], S.SCENE_CAMERA["direction"], 90.0)

    def allreduce(ctx):
        t = torch.from_numpy(ctx.read("Illumination"))
        dist.all_reduce(t)
        ctx.write("Illumination", t.numpy())

    for f in range(2):
        vpt_shard.render_sharded(o, cam, cam, f, rank, world, allreduce)
    sharded = o.read("Illumination")
    # every rank rendered disjoint samples; only rank 0 owns the G-buffer (sample 0)
    owns = bool((o.read("Depth") != 0).any())
    if rank == 0:
        full = common.setup(O.Oracle(W, H), inp, spp=spp, total=3, diffuse=1)
        for f in range(2):
            full.render(cam, cam, f)
        ref = full.read("Illumination")
        q.put(("result", float(np.abs(sharded[..., :3] - ref[..., :3]).max()), bool(np.array_equal(sharded[..., 3], ref[..., 3])), owns))
    else:
        q.put(("other", owns))
    dist.barrier()
    dist.destroy_process_group()


def test_spp_sharding_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res = [g for g in got if g[0] == "result"][0]
    oth = [g for g in got if g[0] == "other"][0]
    assert res[1] < 1e-5, res          # sum of shards / spp == full render (fp32 summation order only)
    assert res[2] and res[3]           # depth channel carried by the sample-0 shard; rank 0 owns the G-buffer
    assert oth[1] is False             # rank 1 never writes the G-buffer


def _worker_local(rank, world, port, q):
    """Rank-local owner mode (vpt_render_shard_local): every rank owns a G-buffer and runs its own ReSTIR pass."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "real-time-path-tracing-voxel-blocks_b200", "python"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import common
    import oracle as O
    import vpt_scenes as S
    import vpt_shard
    O.set_threads(2)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    W, H, spp = 96, 64, 4
    inp = common.scene_inputs((2, 1, 2), noise_fn=O.perlin_noise_chunks, alias_fn=O.build_alias_table)
    o = common.setup(O.Oracle(W, H), inp, spp=spp, total=3, diffuse=1)
    cam = O.camera_from_scene(W, H, S.SCENE_CAMERA["position"], S.SCENE_CAMERA["direction"], 90.0)
    own = []

    def allreduce(ctx):
        own.append(ctx.read("Illumination").copy())
        t = torch.from_numpy(ctx.read("Illumination"))
        dist.all_reduce(t)
        ctx.write("Illumination", t.numpy())

    for f in range(2):
        vpt_shard.render_sharded(o, cam, cam, f, rank, world, allreduce, local_owner=True)
    summed = o.read("Illumination")
    owns = bool((o.read("Depth") != 0).any())
    res = o.read_reservoirs(1)
    valid_res = float((res["M"] > 0).mean())
    if rank == 0:
        # the same two shards rendered one after the other in this process must add up to the all-reduced image
        parts = []
        for r in range(world):
            ctx = common.setup(O.Oracle(W, H), inp, spp=spp, total=3, diffuse=1)
            for f in range(2):
                ctx.render_shard_local(cam, cam, f, r, world)
            parts.append(ctx.read("Illumination"))
        ref = (parts[0][..., :3] + parts[1][..., :3]) / np.float32(spp)
        q.put(("result", float(np.abs(summed[..., :3] - ref).max()), owns, valid_res))
    else:
        q.put(("other", owns, valid_res))
    dist.barrier()
    dist.destroy_process_group()


def test_spp_sharding_local_owner_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + os.getpid() % 200
    procs = [ctx.Process(target=_worker_local, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    res = [g for g in got if g[0] == "result"][0]
    other = [g for g in got if g[0] == "other"][0]
    assert res[1] <= 1e-5, res                      # all-reduced sum == the two shards added up (fp32 summation order)
    assert res[2] and other[1]                      # BOTH ranks own a G-buffer
    assert res[3] > 0.3 and other[2] > 0.3          # ... and keep ReSTIR reservoirs of their own


def _worker_balanced(rank, world, port, q):
    """Cost-balanced contiguous sample ranges (vpt_shard.render_balanced, the cfg4 / cfg5 strong-scaling split)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "real-time-path-tracing-voxel-blocks_b200", "python"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import common
    import oracle as O
    import vpt_scenes as S
    import vpt_shard
    O.set_threads(2)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    W, H, spp = 96, 64, 6
    inp = common.scene_inputs((2, 1, 2), noise_fn=O.perlin_noise_chunks, alias_fn=O.build_alias_table)
    o = common.setup(O.Oracle(W, H), inp, spp=spp, total=3, diffuse=1)
    cam = O.camera_from_scene(W, H, S.SCENE_CAMERA["position"], S.SCENE_CAMERA["direction"], 90.0)

    def allreduce(ctx):
        t = torch.from_numpy(ctx.read("Illumination"))
        dist.all_reduce(t)
        ctx.write("Illumination", t.numpy())

    for f in range(2):
        vpt_shard.render_balanced(o, cam, cam, f, rank, world, spp, allreduce)
    got = o.read("Illumination")
    if rank == 0:
        full = common.setup(O.Oracle(W, H), inp, spp=spp, total=3, diffuse=1)
        for f in range(2):
            full.render(cam, cam, f)
        ref = full.read("Illumination")
        q.put(("result", float(np.abs(got[..., :3] - ref[..., :3]).max()), bool(np.array_equal(got[..., 3], ref[..., 3])), vpt_shard.balanced_ranges(spp, world)))
    dist.barrier()
    dist.destroy_process_group()


def test_balanced_sample_ranges_gloo_world2():
    sys.path.insert(0, os.path.join(ROOT, "real-time-path-tracing-voxel-blocks_b200", "python"))
    import vpt_shard
    for spp, n in ((64, 8), (64, 4), (64, 2), (256, 8), (4, 8), (8, 8), (1, 3), (5, 1)):
        r = vpt_shard.balanced_ranges(spp, n)
        assert len(r) == n and r[0][0] == 0 and sum(c for _, c in r) == spp
        assert all(r[i][0] + r[i][1] == r[i + 1][0] for i in range(n - 1))
        assert r[0][1] >= 1                                     # rank 0 always renders sample 0 (the G-buffer / ReSTIR sample)
        if n > 1 and spp >= 4 * n:
            assert r[0][1] < max(c for _, c in r[1:])           # ... and fewer plain samples than the others
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30000 + os.getpid() % 200
    procs = [ctx.Process(target=_worker_balanced, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[1] < 1e-5 and res[2], res     # the two ranges add up to the full render (fp32 summation order); depth from the sample-0 rank
    assert res[3][0][1] + res[3][1][1] == 6
