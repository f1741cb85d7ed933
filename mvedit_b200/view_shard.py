"""Sharding of the hot loop across the GPUs of one box.  The reference is single-process / single-GPU
(/root/reference/app.py:62-63; SURVEY.md §2.3: no collective on the inference path), so this has no reference counterpart:

  * denoise (UNet / ControlNet / VAE decode) and render are per-view -> rank r owns views [r*N/G, (r+1)*N/G); the ONE data-path
    collective of that part is ``gather_views``: a single ``all_gather_into_tensor`` per step of the decoded targets, images and masks
    packed as one bf16 [n, rs, rs, 4] tensor (the VAE output is bf16 already), shard sizes derived from ``local_range`` -- no count
    exchange, no host sync;
  * the reconstruction (one shared parameter set, sequential Adam steps) is data-parallel over RAYS: every iteration each rank marches /
    decodes / composites a row strip of the iteration's patches, the per-ray outputs (5 floats per ray) are all-gathered
    (``gather_rays``), the objective runs replicated on the full patches, each rank back-propagates its strip and the flat gradient
    buffer is summed with ONE ``all_reduce`` (``allreduce_flat``, 28.7 MB) before the (replicated, bit-identical) Adam update.
    Patch order and occupancy jitter use rank-independent random streams, so the replicas stay identical without a broadcast;
  * ``broadcast_field`` (one flat broadcast) remains for the replicated mode (``nerf.data_parallel = False``).

Works with any initialised torch.distributed backend (NCCL over NVLink on the GPU box; gloo in the CPU tests); without
torch.distributed everything degenerates to the single-GPU identity.
"""
import torch
import torch.distributed as dist


def _on():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def world():
    return (dist.get_rank(), dist.get_world_size()) if _on() else (0, 1)


def local_range(n_views, rank=None, world_size=None):
    """Contiguous block partition; the first (n % G) ranks get one extra view."""
    r, g = world()
    rank = r if rank is None else rank
    g = g if world_size is None else world_size
    base, extra = divmod(n_views, g)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _all_gather_flat(t):
    """[m, ...] -> [G, m, ...] with one collective (all_gather_into_tensor where the backend has it)."""
    g = dist.get_world_size()
    out = torch.empty((g,) + tuple(t.shape), dtype=t.dtype, device=t.device)
    try:
        dist.all_gather_into_tensor(out.view(-1), t.contiguous().view(-1))
    except (RuntimeError, NotImplementedError):          # gloo builds without the flat variant
        dist.all_gather(list(out.unbind(0)), t.contiguous())
    return out


def gather_views(images, masks=None, n_total=None, pack_dtype=torch.bfloat16):
    """Local decoded targets -> all views in global order, ONE collective.
    images [n_local, h, w, 3] (+ masks [n_local, h, w, 1]) fp32 -> ([n_total, h, w, 3], [n_total, h, w, 1]) fp32.
    Shard sizes follow ``local_range(n_total)``; ragged shards are padded to the largest (static shapes, no count exchange)."""
    if not _on():
        return (images, masks) if masks is not None else images
    r, g = world()
    if n_total is None:                                     # equal shards assumed when the caller does not say otherwise
        n_total = images.shape[0] * g
    counts = [local_range(n_total, k, g)[1] - local_range(n_total, k, g)[0] for k in range(g)]
    assert counts[r] == images.shape[0], 'gather_views: local shard does not match local_range'
    c = 3 + (masks.shape[-1] if masks is not None else 0)
    m = max(counts)
    packed = torch.zeros(m, images.shape[1], images.shape[2], c, dtype=pack_dtype, device=images.device)
    packed[:counts[r], ..., :3] = images
    if masks is not None:
        packed[:counts[r], ..., 3:] = masks
    out = _all_gather_flat(packed)                          # [G, m, h, w, c]
    out = out.reshape(g * m, *out.shape[2:]) if len(set(counts)) == 1 else torch.cat([out[k, :counts[k]] for k in range(g)], dim=0)
    out = out.float()
    return (out[..., :3].contiguous(), out[..., 3:].contiguous()) if masks is not None else out


def gather_rays(t, n_patches=1):
    """Per-ray renderer outputs of this rank's row strips [P * rows_local * ps, C] -> the full patches [P * ps * ps, C]
    (patch-major, rows in order).  One collective."""
    if not _on():
        return t
    g = dist.get_world_size()
    out = _all_gather_flat(t)                               # [G, P*rl*ps, C]
    if n_patches == 1:
        return out.reshape(-1, t.shape[-1])
    return out.reshape(g, n_patches, -1, t.shape[-1]).transpose(0, 1).reshape(-1, t.shape[-1])


def allreduce_flat(flat):
    """Sum of the flat gradient buffer over ranks, in place (one collective)."""
    if _on():
        dist.all_reduce(flat)
    return flat


def allreduce_grads(params):
    """Fallback for optimizers without a flat buffer: one all_reduce per parameter gradient."""
    if _on():
        for p in params:
            if p.grad is not None:
                dist.all_reduce(p.grad)


def broadcast_patch_order(batches, src=0):
    """Every rank must walk the same patch permutation: rank ``src``'s draw wins."""
    if not _on():
        return batches
    perm = torch.cat(list(batches), dim=1).contiguous()
    dist.broadcast(perm, src)
    return perm.split(batches[0].shape[1], dim=1)


def broadcast_field(decoder, density_grid, density_bitfield, src=0):
    """Replicated (non data-parallel) reconstruction: make every rank's field identical to rank ``src``'s with ONE broadcast of a
    flat byte buffer (parameters fp32 + occupancy grid fp16 + bitfield u8)."""
    if not _on():
        return
    parts = [p.data for p in decoder.parameters()] + [density_grid, density_bitfield]
    flat = torch.cat([p.reshape(-1).view(torch.uint8) for p in parts])
    dist.broadcast(flat, src)
    o = 0
    for p in parts:
        nb = p.numel() * p.element_size()
        p.copy_(flat[o:o + nb].view(p.dtype).view(p.shape))
        o += nb
