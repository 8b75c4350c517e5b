"""Tile-batch driver: the one data-parallel axis of the labelling path.

Tiles are independent (every processor is tile-local: AHN raster per tile, BGT query per
tile bbox, octree / neighbour grid per call -- reference pipeline.py:186-194 loops over
them serially), so a batch is sharded tile-wise over one process per GPU with NO
collective on the hot path.  The only exchange is the end-of-batch gather of the per-tile
label histograms (the statistic the reference logs per file, analysis_tools.py:8-18):
one ``all_gather`` over ``torch.distributed`` -- NCCL over NVLink/NVSwitch on GPUs, gloo
on CPU for the unit tests.  At <= 512 tiles x 256 bins x 8 B the payload is ~1 MB.
"""
import torch
import torch.distributed as dist


def shard_tiles(n_tiles, rank, world_size):
    """Static round-robin assignment: rank r owns tiles {i : i mod world_size == r}."""
    return list(range(rank, n_tiles, world_size))


def gather_histograms(local_ids, local_hists, n_tiles, device=None):
    """All-gather per-tile histograms.

    local_ids   : list[int] tile indices processed by this rank
    local_hists : int64 tensor [len(local_ids), 256] on ``device``
    Returns an int64 tensor [n_tiles, 256] (on ``device``) identical on every rank, row i
    holding the histogram of tile i.
    """
    if device is None:
        device = local_hists.device
    bins = local_hists.shape[1] if local_hists.ndim == 2 else 256
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = torch.zeros((n_tiles, bins), dtype=torch.int64, device=device)
        if len(local_ids):
            out[torch.as_tensor(local_ids, device=device)] = local_hists.to(device)
        return out
    world = dist.get_world_size()
    per_rank = (n_tiles + world - 1) // world
    # fixed-size payload: [per_rank, 1 + bins] with the tile id in column 0 (-1 = padding)
    payload = torch.full((per_rank, 1 + bins), -1, dtype=torch.int64, device=device)
    k = len(local_ids)
    if k:
        payload[:k, 0] = torch.as_tensor(local_ids, dtype=torch.int64, device=device)
        payload[:k, 1:] = local_hists.to(device)
    gathered = torch.empty((world * per_rank, 1 + bins), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(gathered, payload)
    out = torch.zeros((n_tiles, bins), dtype=torch.int64, device=device)
    valid = gathered[:, 0] >= 0
    out[gathered[valid, 0]] = gathered[valid, 1:]
    return out


def run_batch(n_tiles, process_tile, device, rank=None, world_size=None, process_tiles=None, on_processed=None):
    """Process this rank's share of ``n_tiles`` tiles and gather the label statistics.

    process_tile(tile_index) -> int64 tensor [256] on ``device`` (label histogram of the
    labelled tile), one tile after the other; or
    process_tiles(list of tile indices) -> int64 tensor [len, 256]: the whole share at once, which
    lets the caller keep several tiles in flight (``Pipeline.process_tiles_device(..., workers=W)``).
    ``on_processed()`` is called between the labelling and the gather (bench.py stops its clock there:
    the gather is the end-of-batch exchange, not part of a step).
    Returns (hist [n_tiles, 256], local tile ids).
    """
    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    ids = shard_tiles(n_tiles, rank, world_size)
    if process_tiles is not None:
        local = process_tiles(ids) if ids else torch.zeros((0, 256), dtype=torch.int64, device=device)
    else:
        hists = [process_tile(i) for i in ids]
        local = torch.stack(hists) if hists else torch.zeros((0, 256), dtype=torch.int64, device=device)
    if on_processed is not None:
        on_processed()
    return gather_histograms(ids, local, n_tiles, device), ids
