"""Multi-GPU sharding helpers: one process per GPU, torch.distributed for plumbing.

Two ways the path shards, neither needs a data-path collective while rendering:
  * independent canvases / frames: rank r takes items r, r + world, ... (`my_items`);
  * scanline bands of one large canvas: rank r owns rows [band(r)) of the image and renders them
    with `cb200_canvas_create_band`; the finished RGBA8 bands are gathered once per image with
    all_gather (NCCL over NVLink on GPUs, gloo on CPU for tests).
"""
import torch
import torch.distributed as dist


def band(height, rank, world):
    """Rows [y0, y0 + rows) owned by `rank`: contiguous, exhaustive, sizes differ by at most one."""
    y0 = rank * height // world
    y1 = (rank + 1) * height // world
    return y0, y1 - y0


def my_items(n_items, rank, world):
    """Indices of the independent canvases rank `rank` renders (round robin)."""
    return list(range(rank, n_items, world))


def gather_bands(local_band, height, width):
    """all_gather the per-rank RGBA8 bands (uint8 tensor [rows, width, 4]) into one [height, width, 4]
    image on every rank.  Bands may differ by one row, so gather padded bands and trim."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return local_band
    max_rows = (height + world - 1) // world
    padded = torch.zeros((max_rows, width, 4), dtype=torch.uint8, device=local_band.device)
    padded[:local_band.shape[0]] = local_band
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    rows = [band(height, r, world)[1] for r in range(world)]
    return torch.cat([p[:n] for p, n in zip(parts, rows)], dim=0)
