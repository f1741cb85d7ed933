"""View sharding across the GPUs of one box.  The reference is single-process / single-GPU
(/root/reference/app.py:62-63; SURVEY.md §2.3: no collective on the inference path), so this has no reference counterpart:

  * denoise (UNet/ControlNet) and render are per-view -> rank r owns views [r*N/G, (r+1)*N/G);
  * the ONLY data-path collective is one all_gather per adapter iteration of the decoded targets (images + masks), because the
    reconstruction needs every view;
  * the reconstruction runs replicated on the gathered targets; rank 0's field (28.7 MB table + MLP + occupancy) is broadcast
    afterwards so replicas stay bit-identical despite atomic-order noise in the hash-grid gradient.

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


def gather_views(t):
    """[n_local, ...] -> [n_total, ...] in global view order (one all_gather; ragged shards are padded to the largest)."""
    if not _on():
        return t
    r, g = world()
    counts = [None] * g
    n = torch.tensor([t.shape[0]], device=t.device, dtype=torch.int64)
    all_n = [torch.zeros_like(n) for _ in range(g)]
    dist.all_gather(all_n, n)
    counts = [int(x) for x in all_n]
    m = max(counts)
    if t.shape[0] < m:
        t = torch.cat([t, t.new_zeros(m - t.shape[0], *t.shape[1:])], dim=0)
    out = [torch.empty_like(t) for _ in range(g)]
    dist.all_gather(out, t.contiguous())
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)


def broadcast_field(decoder, density_grid, density_bitfield, src=0):
    """After the replicated reconstruction: make every rank's field identical to rank `src`'s."""
    if not _on():
        return
    for p in decoder.parameters():
        dist.broadcast(p.data, src)
    dist.broadcast(density_grid, src)
    dist.broadcast(density_bitfield, src)
