"""Multi-GPU sharding of scene pairs: pairs are independent, rank r owns pairs
r, r+W, r+2W, ...; the only collective is one all-gather of the (P_local,4,4) transforms (NCCL over NVLink on
GPUs, gloo in the CPU tests).  The reference has no inference-time parallelism (test.py:146-161 is a
single-process loop), so this replaces nothing and adds one collective per batch."""
import torch
import torch.distributed as dist


def shard_pairs(n_pairs, rank, world):
    """Indices of the pairs rank `rank` processes (static round-robin: all pairs cost the same)."""
    return list(range(rank, n_pairs, world))


def local_count(n_pairs, rank, world):
    return len(range(rank, n_pairs, world))


def gather_transforms(local_transforms, n_pairs, rank=None, world=None):
    """All-gather per-rank (P_local,4,4) transforms and re-interleave them into pair order -> (n_pairs,4,4).

    Every rank pads to ceil(n_pairs / world) rows so that one fixed-size all_gather_into_tensor suffices."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world == 1:
        return local_transforms
    per = (n_pairs + world - 1) // world
    buf = torch.zeros((per, 4, 4), dtype=local_transforms.dtype, device=local_transforms.device)
    buf[: local_transforms.shape[0]] = local_transforms
    out = torch.empty((world * per, 4, 4), dtype=buf.dtype, device=buf.device)
    dist.all_gather_into_tensor(out, buf)
    # out[r*per + i] is pair r + i*world
    return out.view(world, per, 4, 4).transpose(0, 1).reshape(world * per, 4, 4)[:n_pairs].contiguous()


def default_workers():
    """Worker threads `register_pairs(workers=...)` uses when asked for "as many as pay off" (env GAUSSREG_WORKERS)."""
    import os
    return max(1, int(os.environ.get("GAUSSREG_WORKERS", "4")))


def register_pairs(model, pairs, cfg=None, neighbor_limits=None, streams=1, pyramid_batch=1, workers=1, distributed=True):
    """Coarse-register a list of scene pairs (BASELINE configs 3 / 4): every rank runs the pairs it owns through
    the single-pair forward (the reference model is batch-1 only, model.py:77-89; pairs never interact), one
    all-gather returns all transforms in pair order to every rank.

    `pairs`: sequence of dicts with ref_points / src_points / ref_feats / src_feats (numpy or tensors), indexed
    globally; each rank touches only `pairs[rank::world]`.  Returns a (len(pairs), 4, 4) float32 CUDA tensor.

    `pyramid_batch` > 1 builds ONE neighbour pyramid per that many pairs (data.precompute_pairs_stack_mode): the
    ~170 latency-bound launches and the host syncs of the pyramid are shared by the whole group, the network still
    runs pair by pair; results are bit-identical.

    `streams` > 1 software-pipelines consecutive pairs over that many CUDA streams: the neighbour pyramid of pair
    i+1 (a few CTAs per kernel, two host syncs for its data-dependent sizes) then overlaps the network of pair i.
    Every per-call scratch buffer of the library is keyed by stream, so the results are bit-identical to the
    sequential order; the first pair always runs alone (it packs the static weights other streams then read).

    `workers` > 1 is the throughput mode of BASELINE configs 3 / 4: that many host threads, each with its own CUDA
    stream, take pairs round-robin.  A single pair's forward is a chain of several hundred mostly small kernels
    that leave most of the 148 SMs idle; pairs on different streams fill them.  The C-ABI calls release the GIL, so
    the threads issue concurrently; the per-stream workspaces keep the pairs independent and every transform is
    bit-identical to the sequential result.

    `pairs` may hold None at the positions other ranks own.  `distributed=False` skips the all-gather (warm-up)."""
    from .config import make_cfg, NEIGHBOR_LIMITS
    from .data import registration_collate_fn_stack_mode
    cfg = cfg or make_cfg()
    limits = neighbor_limits or NEIGHBOR_LIMITS
    use_dist = distributed and dist.is_initialized()
    world = dist.get_world_size() if use_dist else 1
    rank = dist.get_rank() if use_dist else 0
    dev = torch.device("cuda", torch.cuda.current_device())
    mine = shard_pairs(len(pairs), rank, world)
    local = torch.empty((len(mine), 4, 4), dtype=torch.float32, device=dev)
    keys = ("ref_points", "src_points", "ref_feats", "src_feats")

    def one(j, i):
        data = registration_collate_fn_stack_mode([{k: pairs[i][k] for k in keys}], cfg.backbone.num_stages,
                                                  cfg.backbone.init_voxel_size, cfg.backbone.init_radius, limits)
        local[j] = model(data)["estimated_transform"]

    if pyramid_batch > 1:
        from .data import precompute_pairs_stack_mode
        for g0 in range(0, len(mine), pyramid_batch):
            group = mine[g0:g0 + pyramid_batch]
            pyr = precompute_pairs_stack_mode([pairs[i]["ref_points"] for i in group], [pairs[i]["src_points"] for i in group],
                                              cfg.backbone.num_stages, cfg.backbone.init_voxel_size, cfg.backbone.init_radius, limits)
            for j, (i, data) in enumerate(zip(group, pyr)):
                data["features"] = torch.cat([torch.as_tensor(pairs[i]["ref_feats"]), torch.as_tensor(pairs[i]["src_feats"])],
                                             dim=0).to(dev, torch.float32, non_blocking=True)
                data["batch_size"] = 1
                local[g0 + j] = model(data)["estimated_transform"]
    elif workers > 1 and len(mine) > 1:
        import threading
        cur = torch.cuda.current_stream()
        one(0, mine[0])  # alone: packs the static weights every stream then reads
        cur.synchronize()
        pool = _stream_pool(dev, workers)
        errors = []

        def work(w):
            try:
                torch.cuda.set_device(dev)
                with torch.cuda.stream(pool[w]):
                    for j in range(1 + w, len(mine), workers):
                        one(j, mine[j])
                    pool[w].synchronize()
            except BaseException as ex:  # surfaced on the caller's thread below
                errors.append(ex)

        threads = [threading.Thread(target=work, args=(w,), daemon=True) for w in range(workers)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
    elif streams <= 1 or len(mine) <= 1:
        for j, i in enumerate(mine):
            one(j, i)
    else:
        cur = torch.cuda.current_stream()
        one(0, mine[0])
        cur.synchronize()
        pool = _stream_pool(dev, streams)
        for s in pool:
            s.wait_stream(cur)
        for j, i in enumerate(mine[1:], start=1):
            with torch.cuda.stream(pool[j % streams]):
                one(j, i)
        for s in pool:
            cur.wait_stream(s)
    return gather_transforms(local, len(pairs), rank, world)


_STREAM_POOLS = {}


def _stream_pool(dev, n):
    """Streams are kept for the life of the process: the library's per-stream workspaces are keyed by them."""
    pool = _STREAM_POOLS.setdefault(dev.index, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=dev))
    return pool[:n]
