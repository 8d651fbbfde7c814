"""Host-side sharding logic of the multi-GPU paths. Pure Python so the CPU test-suite can exercise it
with gloo and the oracle as the compute stand-in; the GPU path uses the same functions with libvpt + NCCL.

  * spp sharding: rank r of n renders sample indices {k : k mod n == r} of all pixels (vpt_render_shard(begin=r, step=n)),
    the un-normalised fp32 sums are all-reduced (sum) and every rank divides by the total spp (vpt_resolve).
    The G-buffer, depth and the ReSTIR reservoir belong to sample 0, i.e. to rank 0.
  * row bands: rank r owns rows [b[r], b[r+1]) with every boundary a multiple of 4 (FireflyBoilingFilter's 8x4 tile
    statistics must not straddle ranks, FireflyFilter.h:51-65)."""


def sample_shard(rank, nranks):
    """(sample_begin, sample_step) for vpt_render_shard."""
    if not (0 <= rank < nranks):
        raise ValueError("rank out of range")
    return rank, nranks


def samples_of(rank, nranks, spp):
    return list(range(rank, spp, nranks))


def row_bands(height, nranks):
    """nranks+1 row boundaries, multiples of 4 (the last one is `height`)."""
    if nranks < 1 or height < 4 * nranks:
        raise ValueError("need at least 4 rows per rank")
    b = [(height * r // nranks) // 4 * 4 for r in range(nranks)] + [height]
    return b


def render_sharded(ctx, cam, prev_cam, iteration_index, rank, nranks, allreduce_sum, local_owner=False):
    """One spp-sharded frame: ctx is a vpt.Vpt or an oracle.Oracle; allreduce_sum(ctx) sums Illumination over ranks.
    local_owner: every rank's first sample owns that rank's G-buffer / reservoirs / ReSTIR pass (equal work per rank; the sum is
    an average of independent frames instead of one frame of nranks x spp samples with a single ReSTIR sample)."""
    begin, step = sample_shard(rank, nranks)
    if local_owner:
        ctx.render_shard_local(cam, prev_cam, iteration_index, begin, step)
    else:
        ctx.render_shard(cam, prev_cam, iteration_index, begin, step)
    allreduce_sum(ctx)
    ctx.resolve()


def frame_banded(ctx, params, cam, prev_cam, frame, rank, nranks, allreduce_sum, height, gather=True):
    """One frame of the composed multi-GPU path (both rows): every rank renders its samples with a rank-local owner
    (its own G-buffer, reservoirs and ReSTIR pass), the accumulation buffers are summed over the ranks, and every rank denoises its
    row band (vpt_denoise_band: no exchange inside the chain, one history exchange per frame); the bands of the output are
    collected on rank 0."""
    begin, step = sample_shard(rank, nranks)
    ctx.render_shard_local(cam, prev_cam, frame, begin, step)
    allreduce_sum(ctx)
    ctx.resolve()
    b = row_bands(height, nranks)
    ctx.denoise_band(params, cam, prev_cam, frame, frame + 1, b[rank], b[rank + 1])
    if gather and nranks > 1:
        ctx.comm_gather_output(0)


def balanced_ranges(spp, nranks, owner_extra=2.3):
    """Contiguous sample ranges [(begin, count)] per rank for ONE image of `spp` samples (strong scaling, cfg4 / cfg5), sized by cost:
    rank 0 renders sample 0, whose temporal ReSTIR pass is worth ~1.5 plain samples, and denoises the frame (~0.8 at 4K / 64 spp), so
    it gets `owner_extra` fewer plain samples than the others. Every sample is rendered exactly once whatever the split."""
    if nranks == 1:
        return [(0, spp)]
    per = (spp + owner_extra) / nranks                      # cost units per rank
    first = max(1, min(spp, int(round(per - owner_extra))))
    rest = spp - first
    counts = [first] + [rest // (nranks - 1) + (1 if i < rest % (nranks - 1) else 0) for i in range(nranks - 1)]
    out, b = [], 0
    for c in counts:
        out.append((b, c))
        b += c
    return out


def render_balanced(ctx, cam, prev_cam, iteration_index, rank, nranks, spp, allreduce_sum, owner_extra=2.3):
    """Strong-scaling frame: cost-balanced contiguous sample ranges, sum over ranks, resolve."""
    begin, count = balanced_ranges(spp, nranks, owner_extra)[rank]
    ctx.render_range(cam, prev_cam, iteration_index, begin, count)
    allreduce_sum(ctx)
    ctx.resolve()
