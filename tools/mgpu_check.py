#!/usr/bin/env python3
"""Multi-GPU correctness check (run under torchrun, one rank per GPU):
  1. spp-sharded render: rank r renders samples r, r+N, ...; ncclAllReduce(sum) of the accumulation buffer;
     the resolved image must equal the single-GPU render of the same total spp up to fp32 summation order.
  2. row-band sharded denoiser (config 3 shape): every rank denoises its band with NCCL halo exchange; the
     gathered bands must equal the single-GPU chain.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "real-time-path-tracing-voxel-blocks_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
import vpt  # noqa: E402
import vpt_scenes as S  # noqa: E402
import vpt_shard  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    def fresh_uid():  # one ncclUniqueId per communicator
        t = torch.from_numpy(vpt.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8)).cuda()
        dist.broadcast(t, 0)
        return t.cpu().numpy()

    uid = fresh_uid()
    ok = True

    # ---- 1. spp sharding
    W, H, spp = 640, 360, 4 * world
    inp = common.scene_inputs((4, 1, 4))
    g = common.setup(vpt.Vpt(W, H, local), inp, spp=spp, total=3, diffuse=1)
    g.comm_init(rank, world, uid)
    cam = common.scene_camera(W, H, (4, 1, 4))
    for f in range(2):
        vpt_shard.render_sharded(g, cam, cam, f, rank, world, lambda c: c.comm_allreduce_illumination())
    sharded = g.read("Illumination")
    if rank == 0:
        ref = common.setup(vpt.Vpt(W, H, local), inp, spp=spp, total=3, diffuse=1)
        for f in range(2):
            ref.render(cam, cam, f)
        full = ref.read("Illumination")
        m, outl, dmax = common.rel_err_stats(sharded[..., :3], full[..., :3])
        same_w = np.array_equal(sharded[..., 3], full[..., 3])
        print("[mgpu] spp-sharded x%d vs single GPU: mean rel %.3e, outliers %.3e, max abs %.3e, depth channel identical: %s" % (world, m, outl, dmax, same_w))
        ok &= m < 1e-5 and same_w
        ref.close()
    g.close()

    # ---- 2. row-band denoiser
    W, H = 960, 544
    p = S.default_denoising_params()
    cam = vpt.camera_init(W, H)
    cam[6:9] = (0.0, 6.0, 0.0)
    cam = vpt.camera_set_yaw_pitch(cam, 0.0, 0.0)
    rows = vpt_shard.row_bands(H, world)
    r0, r1 = rows[rank], rows[rank + 1]
    b = vpt.Vpt(W, H, local)
    b.comm_init(rank, world, fresh_uid())
    full = vpt.Vpt(W, H, local) if rank == 0 else None
    for f in range(6):
        gb = S.synthetic_gbuffer(W, H, f)
        b.begin_external_frame()
        for name in ("Illumination", "Depth", "NormalRoughness", "Material", "Albedo"):
            b.write(name, gb[name])
        b.denoise_band(p, cam, cam, f, f + 1, r0, r1)
        if full is not None:
            full.begin_external_frame()
            for name in ("Illumination", "Depth", "NormalRoughness", "Material", "Albedo"):
                full.write(name, gb[name])
            full.denoise(p, cam, cam, f, f + 1)
    b.comm_gather_output(0)
    if rank == 0:
        refo = full.read("IlluminationOutput")
        got = b.read("IlluminationOutput")
        m, outl, dmax = common.rel_err_stats(got, refo)
        print("[mgpu] row-band denoiser x%d vs single GPU: mean rel %.3e, outliers %.3e, max abs %.3e, bit-identical: %s"
              % (world, m, outl, dmax, np.array_equal(got, refo)))
        ok &= m < 1e-6
    b.close()
    if full is not None:
        full.close()

    # ---- 3. band denoiser on rendered frames with a MOVING camera (reprojected history crosses the band boundaries): every rank
    # renders the same frames itself (no spp sharding here), denoises its band; rank 0 also runs the whole chain on one GPU
    W, H = 960, 544
    inp = common.scene_inputs((4, 1, 4))
    b = common.setup(vpt.Vpt(W, H, local), inp, spp=1, total=3, diffuse=1)
    b.comm_init(rank, world, fresh_uid())
    full = common.setup(vpt.Vpt(W, H, local), inp, spp=1, total=3, diffuse=1) if rank == 0 else None
    cam = common.scene_camera(W, H, (4, 1, 4))
    prev = cam
    for f in range(6):
        if f >= 2:
            cam = vpt.camera_set_yaw_pitch(prev, prev[15] + np.float32(0.3 * np.pi / 180.0), prev[16])
        b.render(cam, prev, f)
        b.denoise_band(p, cam, prev, f, f + 1, r0, r1)
        if full is not None:
            full.render(cam, prev, f)
            full.denoise(p, cam, prev, f, f + 1)
        prev = cam
    b.comm_gather_output(0)
    if rank == 0:
        refo, got = full.read("IlluminationOutput"), b.read("IlluminationOutput")
        same_hist = np.array_equal(b.read("PrevIllumination")[r0:r1], full.read("PrevIllumination")[r0:r1])
        print("[mgpu] band denoiser x%d on rendered frames, moving camera: output bit-identical: %s, own-band history bit-identical: %s"
              % (world, np.array_equal(got, refo), same_hist))
        ok &= np.array_equal(got, refo) and same_hist
    okt = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(okt, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(okt.item()) else 1)


if __name__ == "__main__":
    main()
