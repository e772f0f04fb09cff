"""Neighbour pyramid + collate on the GPU (reference: geotransformer/utils/data.py:13-77,139-189).

Same function names, arguments and returned dict as the reference, but every tensor lives on the
CUDA device and the 4 grid subsamples + 13 radius searches are gaussreg_b200 kernels.  Host syncs:
one for the stage lengths, one for the 13 neighbour-table widths (the reference's tensors have
data-dependent shapes)."""
import ctypes
import os

import numpy as np
import torch

from . import _lib, ext


class LazyTables(list):
    """List of neighbour tables that is filled on first access.

    The table WIDTHS are data-dependent (`min(max_count, limit)`, radius_search.py:24-27), so trimming the limit-wide
    device tables needs one device->host read.  Deferring that read to the first access lets the caller queue
    coordinate-only GPU work (point-to-node partition, structure embedding) behind the radius searches first: the
    host then waits for the widths while the GPU is busy instead of idle.  One `finalize` is shared by the three
    lists of a pyramid; any list operation triggers it."""

    def __init__(self, finalize):
        super().__init__()
        self._finalize = finalize

    def _ensure(self):
        f = self._finalize
        if f is not None:
            f()

    def __getitem__(self, i):
        self._ensure()
        return list.__getitem__(self, i)

    def __len__(self):
        self._ensure()
        return list.__len__(self)

    def __iter__(self):
        self._ensure()
        return list.__iter__(self)

    def __repr__(self):
        self._ensure()
        return list.__repr__(self)


def precompute_data_stack_mode(points, lengths, num_stages, voxel_size, radius, neighbor_limits, lazy=True, early=None,
                               features=None, after_chain=None):
    """utils/data.py:13-77 on the GPU.  Returns the reference's dict plus `lengths_host` (python ints per stage and
    cloud, read in the same device->host transfer as the stage sizes, so that the model needs no sync of its own).
    With `lazy` (default) the three table lists are LazyTables: same contents, trimmed on first access.

    Issue order: the four grid subsamples form a dependent chain whose long kernels (the hash-order replay) occupy a
    handful of SMs; every kernel of the pyramid takes its true sizes from device memory (the host only knows upper
    bounds), so all 13 radius searches are queued on a helper stream behind per-stage events and run UNDERNEATH the
    chain.  The host then reads the stage sizes as soon as the chain is done, and the table widths lazily."""
    assert num_stages == len(neighbor_limits)
    dev = points.device if points.is_cuda else ext._device()
    points = points.to(dev, torch.float32).contiguous()
    lengths = lengths.to(dev, torch.int64).contiguous()
    n0 = points.shape[0]
    nb = lengths.shape[0]
    main = torch.cuda.current_stream(dev)
    side = ext.side_stream(dev) if os.environ.get("GAUSSREG_PYRAMID_STREAMS", "2") != "1" else None

    # the searches' row-count maxima accumulate into this (zero-filled BEFORE the helper stream forks off the caller's)
    counts_dev = torch.zeros((3 * num_stages - 2,), dtype=torch.int32, device=dev)
    # GAUSSREG_CHAIN_PRIORITY=1: the subsampling chain runs on a high-priority stream of its own (ext.chain_stream), so that
    # its small kernels go ahead of the wide search kernels of the helper stream instead of queueing behind their CTAs (the
    # two small-stage calls: 0.17 / 0.19 ms next to the searches, 0.09 / 0.09 ms with priority).  Every buffer the chain
    # writes is allocated HERE, on the caller's stream / allocator pool, before the other streams fork off.
    native = os.environ.get("GAUSSREG_PYRAMID_NATIVE", "1") != "0"
    chain = ext.chain_stream(dev) if side is not None and os.environ.get("GAUSSREG_CHAIN_PRIORITY", "0") == "1" else None
    native_chain = native and side is not None and num_stages > 1 and os.environ.get("GAUSSREG_CHAIN_NATIVE", "1") != "0"
    S = num_stages - 1
    if native_chain:
        # two allocations for all outputs of the chain: the stage points, and [totals | lengths of every stage] -- which
        # is exactly the vector the host reads below (no torch.cat)
        fpool = torch.empty((S, n0, 3), dtype=torch.float32, device=dev)
        ipool = torch.empty((S + num_stages * nb,), dtype=torch.int64, device=dev)
        ipool[S:S + nb].copy_(lengths, non_blocking=True)
    elif chain is not None:
        chain_out = [ext.grid_subsample_outputs(n0, nb, dev) for _ in range(1, num_stages)]
    if side is not None:
        # the helper streams start where the caller's stream is NOW: inputs are ready, and every kernel that may still
        # read a recycled buffer (the previous pair's forward) has been ordered before it
        start = torch.cuda.Event()
        start.record(main)
        side.wait_event(start)
        if chain is not None:
            chain.wait_event(start)

    # --- grid subsampling chain, device-side lengths, capacity-sized outputs
    sizes_dev = None
    pts_cap, len_dev, totals, ready = [points], [lengths], [], [None]
    with torch.cuda.stream(chain if chain is not None else main):
        if native_chain:
            # ONE C-ABI call for the whole chain (gr_grid_subsample_chain); the per-stage events come back as raw handles
            # for gr_radius_pyramid
            L = _lib.lib()
            vox = (ctypes.c_float * S)(*[voxel_size * (2 ** i) for i in range(1, num_stages)])
            ev_out = (ctypes.c_void_p * num_stages)()
            ws = ext._workspace(L.gr_grid_subsample_workspace_size(n0, nb), dev)
            st = L.gr_grid_subsample_chain(points.data_ptr(), lengths.data_ptr(), nb, n0, vox, S, fpool.data_ptr(),
                                           ipool.data_ptr() + 8 * (S + nb), ipool.data_ptr(), ws.data_ptr(), ws.numel(), ev_out,
                                           ext._stream())
            _lib.check(st, "grid_subsample_chain")
            for i in range(S):
                pts_cap.append(fpool[i])
                len_dev.append(ipool[S + nb * (i + 1):S + nb * (i + 2)])
                totals.append(None)
                ready.append(ev_out[i + 1])
            sizes_dev = ipool
        else:
            for i in range(1, num_stages):
                voxel_size_i = voxel_size * (2 ** i)
                out, out_len, out_total = ext.grid_subsample_device(pts_cap[-1], len_dev[-1], voxel_size_i, n_points=n0,
                                                                    out=chain_out[i - 1] if chain is not None else None)
                pts_cap.append(out)
                len_dev.append(out_len)
                totals.append(out_total)
                if side is not None:
                    ev = torch.cuda.Event()
                    ev.record(chain if chain is not None else main)
                    ready.append(ev)
        if chain is not None:
            chain_done = torch.cuda.Event()
            chain_done.record(chain)
    if chain is not None:
        main.wait_event(chain_done)  # the caller's stream continues behind the chain
    if sizes_dev is None:
        sizes_dev = torch.cat(totals + len_dev) if totals else torch.cat(len_dev)
    if after_chain is not None:
        after_chain()  # host work of the caller that only has to happen before the network runs (the feature upload)

    # --- buffers (all on the caller's stream / allocator pool, every stage sized by the upper bound n0): allocated AFTER the
    # subsampling chain has been queued, so that the GPU is already working while the host does this (0.1 ms per pair)
    specs = []  # (key, query stage, support stage, radius, limit)
    r = radius
    for i in range(num_stages):
        specs.append(("neighbors", i, i, r, neighbor_limits[i]))
        if i < num_stages - 1:
            specs.append(("subsampling", i + 1, i, r, neighbor_limits[i]))
            specs.append(("upsampling", i, i + 1, r * 2, neighbor_limits[i + 1]))
        r *= 2
    # ONE allocation for the 13 tables and one for the 5 cell grids (18 torch.empty calls were 0.1 ms of host time between the
    # last subsample call and the searches, i.e. in front of the stage-size read the whole step waits for)
    table_off, o = [], 0
    for (_, _, _, _, limit) in specs:
        table_off.append(o)
        o += (n0 * limit + 31) // 32 * 32  # every table starts on a 256-byte boundary, as separate allocations would
    table_pool = torch.empty((max(o, 1),), dtype=torch.int64, device=dev)
    table_base = table_pool.data_ptr()

    def table(j):  # (n0, limit) view of table j; built where it is needed (finalize), not in front of the searches
        return table_pool[table_off[j]:table_off[j] + n0 * specs[j][4]].view(n0, specs[j][4])

    assert len(specs) == counts_dev.shape[0]
    grid_bytes = (max(_lib.lib().gr_radius_neighbors_workspace_size(0, n0, nb), 256) + 255) // 256 * 256  # upper bound n0
    grid_pool = torch.empty((num_stages * grid_bytes,), dtype=torch.uint8, device=dev)
    grids = [grid_pool[i * grid_bytes:(i + 1) * grid_bytes] for i in range(num_stages)]
    # --- radius searches (limit-wide tables, widths come back later).  Stage i's support cloud is searched with radius
    # r_i by "neighbors" and "subsampling" of stage i and (r_i = 2 r_{i-1}) by "upsampling" of stage i-1: one cell grid per
    # stage serves all three (5 grids for 13 searches).
    built = [False] * num_stages

    def searches_native(stream, first=0, last=None, built_mask=0):
        # one C-ABI call for the searches [first, last) (gr_radius_pyramid): ~50 us of host time instead of ~0.6 ms for 13
        L = _lib.lib()
        counts_base = counts_dev.data_ptr()
        sel = specs[first:last]
        arr = (_lib.PyramidSearch * len(sel))()
        for j, (key, qs, ss, rad, limit) in enumerate(sel, start=first):
            a = arr[j - first]
            a.query_stage, a.support_stage, a.radius, a.limit = qs, ss, float(rad), int(limit)
            a.out_idx, a.out_max_count = table_base + 8 * table_off[j], counts_base + 4 * j
        vp = ctypes.c_void_p * num_stages
        pts_arr = vp(*[p.data_ptr() for p in pts_cap])
        len_arr = vp(*[l.data_ptr() for l in len_dev])
        ws_arr = vp(*[g.data_ptr() for g in grids])
        ev_arr = vp(*[(e if e is None or isinstance(e, int) else e.cuda_event) for e in ready]) if side is not None else None
        st = L.gr_radius_pyramid(pts_arr, len_arr, num_stages, nb, n0, ws_arr, min(g.numel() for g in grids), ev_arr, arr, len(sel),
                                 built_mask, stream.cuda_stream)
        _lib.check(st, "radius_pyramid")

    early_out = None
    if early is not None and native and side is not None and features is not None and specs[0][:3] == ("neighbors", 0, 0):
        # Stage 0 needs nothing from the subsampling chain: its neighbour table first, then the caller's stage-0 work
        # (the first two backbone blocks) on the caller's stream, right behind the chain.  The GPU then stays busy while
        # the host waits for the stage sizes below and prepares what depends on them.  The table is passed untrimmed
        # (n0, limit): its padding entries are shadow neighbours.
        with torch.cuda.stream(side):
            searches_native(side, 0, 1)
            ev0 = torch.cuda.Event()
            ev0.record(side)
        main.wait_event(ev0)
        early_out = early(features, pts_cap[0], table(0))

    def searches():
        if native:
            if early_out is not None:
                return searches_native(side, 1, None, built_mask=1)
            return searches_native(side if side is not None else main)
        waited = 0
        for j, (key, qs, ss, rad, limit) in enumerate(specs):
            need = max(qs, ss)
            if side is not None:
                while waited < need:
                    waited += 1
                    side.wait_event(ready[waited])
            ext.radius_neighbors_device(pts_cap[qs], pts_cap[ss], len_dev[qs], len_dev[ss], rad, limit, out=table(j),
                                        grid_ws=grids[ss], reuse_grid=built[ss], max_count=counts_dev[j:j + 1])
            built[ss] = True

    done = None
    if side is not None:
        with torch.cuda.stream(side):
            searches()
            done = torch.cuda.Event()
            done.record(side)
    else:
        searches()

    host = sizes_dev.cpu().tolist()  # sync 1: stage sizes and per-cloud lengths (waits for the subsample chain only)
    tot = host[:len(totals)]
    lengths_host = [host[len(totals) + s * nb: len(totals) + (s + 1) * nb] for s in range(num_stages)]
    sizes = [n0] + [int(t) for t in tot]
    points_list = [pts_cap[i][: sizes[i]] for i in range(num_stages)]
    out = {"points": points_list, "lengths": len_dev, "lengths_host": lengths_host}
    if early_out is not None:
        out["early_features"] = early_out

    def finalize():
        if done is not None:
            # read the widths on the HELPER stream: the copy then completes as soon as the searches do, instead of queueing
            # behind whatever the caller's stream is running by now (the structure embedding) -- the host can trim the
            # tables and issue the backbone while that kernel is still busy
            with torch.cuda.stream(side):
                widths = counts_dev.cpu().tolist()  # sync 2
            torch.cuda.current_stream(dev).wait_event(done)  # the searches' tables become visible to the caller's stream
        else:
            widths = counts_dev.cpu().tolist()  # sync 2
        for j, (w, (key, qs, _, _, limit)) in enumerate(zip(widths, specs)):
            w = min(int(w), limit)
            # rows [0, size of the query stage), columns [0, w) of table j: one view op per table (this loop runs between
            # the embedding's launch and the backbone call, where the host is the bottleneck)
            t = torch.as_strided(table_pool, (sizes[qs], w), (limit, 1), table_off[j])
            list.append(out[key], t if w == limit else t.contiguous())
        for key in ("neighbors", "subsampling", "upsampling"):
            out[key]._finalize = None

    for key in ("neighbors", "subsampling", "upsampling"):
        out[key] = LazyTables(finalize)
    if not lazy:
        finalize()
        for key in ("neighbors", "subsampling", "upsampling"):
            out[key] = list(out[key])
    return out


def registration_collate_fn_stack_mode(data_dicts, num_stages, voxel_size, search_radius, neighbor_limits,
                                       precompute_data=True, early=None):
    """utils/data.py:139-189.  Points are organised [ref_1..ref_B, src_1..src_B].

    `early` (extension, e.g. ``model.backbone.forward_early``): a callable (features, points0, neighbors0) -> tensor that
    is queued on the GPU as soon as stage 0 of the pyramid is known; its result is returned as ``early_features`` and
    picked up by ``KPConvFPN.forward``."""
    batch_size = len(data_dicts)
    collated = {}
    for data_dict in data_dicts:
        for key, value in data_dict.items():
            if isinstance(value, np.ndarray):
                value = torch.from_numpy(value)
            collated.setdefault(key, []).append(value)
    dev = ext._device()
    feat_list = collated.pop("ref_feats") + collated.pop("src_feats")
    points_list = collated.pop("ref_points") + collated.pop("src_points")
    lengths = torch.LongTensor([p.shape[0] for p in points_list])
    main = torch.cuda.current_stream(dev)
    late = precompute_data and (early is None or batch_size != 1) and not any(t.is_cuda for t in feat_list)
    up = {}

    def upload_features():
        # The features are first read by the backbone: their upload runs on a copy stream underneath the subsampling chain,
        # and is issued only after that chain has been queued (the points go first, the pyramid starts from them).  The copy
        # stream starts where the caller's stream was at the beginning of this call (a recycled buffer's previous readers
        # are ordered before it).
        cs = ext.copy_stream(dev)
        cs.wait_event(start)
        with torch.cuda.stream(cs):
            up["feats"] = ext.h2d_rows(feat_list, dev)
            up["ready"] = torch.cuda.Event()
            up["ready"].record(cs)
        up["feats"].record_stream(main)

    if late:
        start = torch.cuda.Event()
        start.record(main)
    else:
        up["feats"] = ext.h2d_rows(feat_list, dev)
    points = ext.h2d_rows(points_list, dev)
    if batch_size == 1:
        for key, value in collated.items():
            collated[key] = value[0]
    if precompute_data:
        collated.update(precompute_data_stack_mode(points, lengths, num_stages, voxel_size, search_radius, neighbor_limits,
                                                   early=early if batch_size == 1 else None, features=up.get("feats"),
                                                   after_chain=upload_features if late else None))
        if late:
            main.wait_event(up["ready"])  # long complete: the host has just waited for the subsampling chain
    else:
        collated["points"] = points
        collated["lengths"] = lengths.to(dev)
    collated["features"] = up["feats"]
    collated["batch_size"] = batch_size
    return collated


def calibrate_neighbors_stack_mode(dataset, collate_fn, num_stages, voxel_size, search_radius, keep_ratio=0.8,
                                   sample_threshold=2000):
    """utils/data.py:192-217: neighbour limits such that `keep_ratio` of the points keep all their neighbours.

    Same signature and integer-exact result; the pyramids come from the CUDA kernels (through `collate_fn`, normally
    `registration_collate_fn_stack_mode`), the neighbourhood-size histograms are accumulated on the device and only
    the (num_stages, hist_n) table crosses to the host."""
    hist_n = int(np.ceil(4 / 3 * np.pi * (search_radius / voxel_size + 1) ** 3))
    max_neighbor_limits = [hist_n] * num_stages
    neighbor_hists = None
    for i in range(len(dataset)):
        data_dict = collate_fn([dataset[i]], num_stages, voxel_size, search_radius, max_neighbor_limits, precompute_data=True)
        hists = []
        for neighbors in data_dict["neighbors"]:
            counts = (neighbors < neighbors.shape[0]).sum(dim=1)
            hists.append(torch.bincount(counts, minlength=hist_n)[:hist_n])
        h = torch.stack(hists)
        neighbor_hists = h if neighbor_hists is None else neighbor_hists + h
        if int(neighbor_hists.sum(dim=1).min()) > sample_threshold:
            break
    neighbor_hists = neighbor_hists.cpu().numpy().astype(np.int32)
    cum_sum = np.cumsum(neighbor_hists.T, axis=0)
    neighbor_limits = np.sum(cum_sum < (keep_ratio * cum_sum[hist_n - 1, :]), axis=0)
    return neighbor_limits


def precompute_pairs_stack_mode(ref_points_list, src_points_list, num_stages, voxel_size, radius, neighbor_limits):
    """ONE neighbour pyramid for P pairs (BASELINE configs 3 / 4), returned as P single-pair pyramids.

    The clouds are stacked [ref_1..ref_P, src_1..src_P] as `registration_collate_fn_stack_mode` stacks a batch
    (utils/data.py:139-189) and go through precompute_data_stack_mode once: every grid subsample / radius search
    kernel then carries P pairs' work per launch.  csrc/pairs.cu re-orders each stage to pair-major order and
    re-bases the neighbour indices, so that pair i's `points / neighbors / subsampling / upsampling` are row
    slices (views) -- bit-identical to what the single-pair call returns, including the table widths
    `min(max_count_of_that_pair, limit)`.  Host syncs: three per batch instead of two per pair.

    Returns a list of dicts with the keys of precompute_data_stack_mode plus `lengths_host` (python ints, lets the
    model skip its own device->host read of the stage lengths)."""
    from . import _lib
    L = _lib.lib()
    P = len(ref_points_list)
    assert P == len(src_points_list) and P > 0
    dev = ext._device()
    clouds = [torch.as_tensor(p) for p in list(ref_points_list) + list(src_points_list)]
    points = torch.cat([c.to(dev, torch.float32, non_blocking=True) for c in clouds], dim=0)
    lengths = torch.tensor([c.shape[0] for c in clouds], dtype=torch.int64, device=dev)
    g = precompute_data_stack_mode(points, lengths, num_stages, voxel_size, radius, neighbor_limits)
    st = ext._stream()
    zero = torch.zeros((1,), dtype=torch.int64, device=dev)
    ref_off = [torch.cat([zero, torch.cumsum(l[:P], 0)]) for l in g["lengths"]]   # (P+1,) per stage
    src_off = [torch.cat([zero, torch.cumsum(l[P:], 0)]) for l in g["lengths"]]
    pm_points = []
    for i in range(num_stages):
        src, out = g["points"][i].contiguous(), torch.empty_like(g["points"][i])
        _lib.check(L.gr_pair_major_rows(src.data_ptr(), 3, src.shape[0], ref_off[i].data_ptr(), src_off[i].data_ptr(), P,
                                        out.data_ptr(), st), "pair_major_rows")
        pm_points.append(out)
    tables, widths = {}, []
    for key, q_stage, s_stage in ([("neighbors", i, i) for i in range(num_stages)] +
                                  [("subsampling", i + 1, i) for i in range(num_stages - 1)] +
                                  [("upsampling", i, i + 1) for i in range(num_stages - 1)]):
        idx = q_stage if key != "subsampling" else s_stage
        t = g[key][idx]
        ld = t.stride(0)
        out = torch.empty((t.shape[0], ld), dtype=torch.int64, device=dev)
        w = torch.empty((P,), dtype=torch.int32, device=dev)
        _lib.check(L.gr_pair_major_table(t.data_ptr(), ld, t.shape[1], t.shape[0], ref_off[q_stage].data_ptr(),
                                         src_off[q_stage].data_ptr(), ref_off[s_stage].data_ptr(), src_off[s_stage].data_ptr(),
                                         P, out.data_ptr(), w.data_ptr(), st), "pair_major_table")
        tables.setdefault(key, []).append(out)
        widths.append(w)
    host_len = torch.stack(g["lengths"]).cpu().tolist()                 # (stages, 2P)     -- sync 3
    host_w = torch.stack(widths).cpu().tolist()                         # (13, P)
    order = ([("neighbors", i) for i in range(num_stages)] + [("subsampling", i) for i in range(num_stages - 1)] +
             [("upsampling", i) for i in range(num_stages - 1)])
    limits = {"neighbors": lambda i: neighbor_limits[i], "subsampling": lambda i: neighbor_limits[i],
              "upsampling": lambda i: neighbor_limits[i + 1]}
    q_stage_of = {"neighbors": lambda i: i, "subsampling": lambda i: i + 1, "upsampling": lambda i: i}
    pm_start = [[0] * (P + 1) for _ in range(num_stages)]
    for s_ in range(num_stages):
        for i in range(P):
            pm_start[s_][i + 1] = pm_start[s_][i] + host_len[s_][i] + host_len[s_][P + i]
    out_list = []
    for i in range(P):
        d = {"points": [], "lengths": [], "lengths_host": [], "neighbors": [], "subsampling": [], "upsampling": []}
        for s_ in range(num_stages):
            a, b = pm_start[s_][i], pm_start[s_][i + 1]
            d["points"].append(pm_points[s_][a:b])
            d["lengths_host"].append((host_len[s_][i], host_len[s_][P + i]))
            d["lengths"].append(torch.stack([g["lengths"][s_][i], g["lengths"][s_][P + i]]))
        for row, (key, k) in enumerate(order):
            qs = q_stage_of[key](k)
            a, b = pm_start[qs][i], pm_start[qs][i + 1]
            w = min(int(host_w[row][i]), limits[key](k))
            d[key].append(tables[key][k][a:b, :w])
        out_list.append(d)
    return out_list
