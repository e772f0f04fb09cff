"""Neighbour pyramid + collate on the GPU (reference: geotransformer/utils/data.py:13-77,139-189).

Same function names, arguments and returned dict as the reference, but every tensor lives on the
CUDA device and the 4 grid subsamples + 13 radius searches are gaussreg_b200 kernels.  Host syncs:
one for the stage lengths, one for the 13 neighbour-table widths (the reference's tensors have
data-dependent shapes)."""
import numpy as np
import torch

from . import ext


def precompute_data_stack_mode(points, lengths, num_stages, voxel_size, radius, neighbor_limits):
    assert num_stages == len(neighbor_limits)
    dev = points.device if points.is_cuda else ext._device()
    points = points.to(dev, torch.float32).contiguous()
    lengths = lengths.to(dev, torch.int64).contiguous()
    n0 = points.shape[0]

    # --- grid subsampling chain, device-side lengths, buffers sized by the upper bound
    pts_cap, len_dev, totals = [points], [lengths], []
    for i in range(1, num_stages):
        voxel_size_i = voxel_size * (2 ** i)
        out, out_len, out_total = ext.grid_subsample_device(pts_cap[-1], len_dev[-1], voxel_size_i,
                                                            n_points=pts_cap[-1].shape[0])
        pts_cap.append(out)
        len_dev.append(out_len)
        totals.append(out_total)
    if totals:
        tot = torch.cat(totals).cpu().tolist()  # sync 1
    else:
        tot = []
    sizes = [n0] + [int(t) for t in tot]
    points_list = [pts_cap[i][: sizes[i]] for i in range(num_stages)]
    lengths_list = len_dev

    # --- radius searches into limit-wide tables, then one sync for the widths
    # Stage i's support cloud is searched with radius r_i by "neighbors" and "subsampling" of stage i and (r_i = 2 r_{i-1})
    # by "upsampling" of stage i-1: one cell grid per stage serves all three (5 grids for 13 searches).
    tables, counts, meta = [], [], []
    grids = [ext.radius_grid_workspace(points_list[i], lengths_list[i]) for i in range(num_stages)]
    built = [False] * num_stages
    r = radius
    for i in range(num_stages):
        cur_p, cur_l = points_list[i], lengths_list[i]
        t, c = ext.radius_neighbors_device(cur_p, cur_p, cur_l, cur_l, r, neighbor_limits[i], grid_ws=grids[i], reuse_grid=built[i])
        built[i] = True
        tables.append(t); counts.append(c); meta.append(("neighbors", neighbor_limits[i]))
        if i < num_stages - 1:
            sub_p, sub_l = points_list[i + 1], lengths_list[i + 1]
            t, c = ext.radius_neighbors_device(sub_p, cur_p, sub_l, cur_l, r, neighbor_limits[i], grid_ws=grids[i], reuse_grid=True)
            tables.append(t); counts.append(c); meta.append(("subsampling", neighbor_limits[i]))
            t, c = ext.radius_neighbors_device(cur_p, sub_p, cur_l, sub_l, r * 2, neighbor_limits[i + 1], grid_ws=grids[i + 1],
                                               reuse_grid=built[i + 1])
            built[i + 1] = True
            tables.append(t); counts.append(c); meta.append(("upsampling", neighbor_limits[i + 1]))
        r *= 2
    widths = torch.cat(counts).cpu().tolist()  # sync 2
    out = {"points": points_list, "lengths": lengths_list, "neighbors": [], "subsampling": [], "upsampling": []}
    for t, w, (key, limit) in zip(tables, widths, meta):
        w = min(int(w), limit)
        out[key].append(t[:, :w] if w == t.shape[1] else t[:, :w].contiguous())
    return out


def registration_collate_fn_stack_mode(data_dicts, num_stages, voxel_size, search_radius, neighbor_limits,
                                       precompute_data=True):
    """utils/data.py:139-189.  Points are organised [ref_1..ref_B, src_1..src_B]."""
    batch_size = len(data_dicts)
    collated = {}
    for data_dict in data_dicts:
        for key, value in data_dict.items():
            if isinstance(value, np.ndarray):
                value = torch.from_numpy(value)
            collated.setdefault(key, []).append(value)
    dev = ext._device()
    feats = torch.cat(collated.pop("ref_feats") + collated.pop("src_feats"), dim=0).to(dev, non_blocking=True)
    points_list = collated.pop("ref_points") + collated.pop("src_points")
    lengths = torch.LongTensor([p.shape[0] for p in points_list])
    points = torch.cat(points_list, dim=0)
    if batch_size == 1:
        for key, value in collated.items():
            collated[key] = value[0]
    collated["features"] = feats
    if precompute_data:
        collated.update(precompute_data_stack_mode(points, lengths, num_stages, voxel_size, search_radius, neighbor_limits))
    else:
        collated["points"] = points.to(dev)
        collated["lengths"] = lengths.to(dev)
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
