"""Row-strip sharding of one very large image over several GPUs (BASELINE.json config 5).

Host-side logic only (numpy + an injected communicator); all pixel work happens in the C-ABI strip entry points:
    msg_meanshift_filter_strip_dev   filter rows [row0,row1) from a buffer holding rows [halo0,halo1) (global coordinates)
    msg_label_strip_dev              union-find labels of a strip, provisional label = 1 + global index of first pixel
    msg_seam_pairs_dev               equivalence pairs across a seam (upper strip's last row vs lower strip's first row)
    msg_apply_label_map_dev          rewrite a strip with the resolved equivalences
The only communication is (1) input halo rows / one boundary row per seam between neighbours (NVLink P2P copies or
NCCL send/recv) and (2) an all-gather of the seam pair lists (NCCL); every rank then solves the same small union-find.
There is no oracle and no CPU pixel path in here.
"""
import numpy as np


def plan_strips(height, n_strips, max_level):
    """Row ranges [(row0,row1), ...]: contiguous, non-empty, every row0 a multiple of 2^max_level (pyramid phase)."""
    a = 1 << max_level
    if n_strips < 1:
        raise ValueError("n_strips must be >= 1")
    units = (height + a - 1) // a
    if units < n_strips:
        raise ValueError("image too small for %d strips at alignment %d" % (n_strips, a))
    bounds = [min(height, ((units * k) // n_strips) * a) for k in range(n_strips)] + [height]
    return [(bounds[k], bounds[k + 1]) for k in range(n_strips)]


def halo_range(row0, row1, height, halo, max_level):
    """Rows a rank must hold to filter [row0,row1): clipped to the image, start aligned to 2^max_level."""
    a = 1 << max_level
    h0 = max(0, row0 - halo)
    h0 -= h0 % a
    h1 = min(height, row1 + halo)
    return h0, h1


def resolve_pairs(pairs):
    """Union-find over seam equivalence pairs.

    pairs: int array (n,2) of global provisional labels (any order, duplicates allowed).
    Returns (from_sorted, to): every label that must change and the label (smallest of its class) it becomes.
    Deterministic: all ranks compute the same map from the same gathered pairs."""
    pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    if len(pairs) == 0:
        return np.zeros(0, np.int32), np.zeros(0, np.int32)
    labels, inv = np.unique(pairs, return_inverse=True)
    inv = inv.reshape(-1, 2)
    n = len(labels)
    try:                                          # C-speed connected components of the pair graph
        from scipy.sparse import coo_matrix
        from scipy.sparse.csgraph import connected_components
        g = coo_matrix((np.ones(len(inv), np.int8), (inv[:, 0], inv[:, 1])), shape=(n, n))
        _, comp = connected_components(g, directed=False)
    except ImportError:                           # same result with a plain union-find
        parent = np.arange(n)

        def find(i):
            r = i
            while parent[r] != r:
                r = parent[r]
            while parent[i] != r:
                parent[i], i = r, parent[i]
            return r

        for a, b in inv:
            ra, rb = find(a), find(b)
            if ra != rb:
                parent[max(ra, rb)] = min(ra, rb)
        comp = np.array([find(i) for i in range(n)])
    # smallest label of every component (labels is sorted ascending, so the first index of a component is its minimum)
    first = np.full(comp.max() + 1, n, np.int64)
    np.minimum.at(first, comp, np.arange(n))
    roots = first[comp]
    changed = roots != np.arange(n)
    return labels[changed].astype(np.int32), labels[roots[changed]].astype(np.int32)


def resolve_dense(quads, n_roots, strips, width):
    """Single-exchange seam resolution + dense numbering tables (msg_seam_quads_dev / msg_strip_finalize_dense_dev).

    quads   : int array (m,4): (A, B, rankA + 1, rankB + 1) over all seams -- provisional labels (1 + global index of the
              strip-local root) of two equivalent regions and the ranks of their roots inside their strips
    n_roots : roots per strip before resolution (msg_strip_rank_dev on the provisional labels)
    strips  : [(row0,row1), ...];  width: image width
    Returns (frm, dense, offsets, frm_lo, total): frm = sorted labels that merge into a smaller label, dense[j] = dense id
    (1..total in raster order of first pixel) of the class frm[j] joins, offsets[s] = surviving roots of the strips above s,
    frm_lo[s] = index of the first entry of frm inside strip s, total = regions of the whole image.
    Deterministic: every rank derives the same tables from the same gathered quads."""
    quads = np.asarray(quads, dtype=np.int64).reshape(-1, 4)
    n_roots = np.asarray(n_roots, dtype=np.int64)
    starts = np.array([r0 * width + 1 for r0, _ in strips], np.int64)          # first label value of every strip
    frm, to = resolve_pairs(quads[:, :2])
    frm64, to64 = frm.astype(np.int64), to.astype(np.int64)
    frm_lo = np.searchsorted(frm64, starts)                                     # entries of frm before each strip
    removed = np.diff(np.concatenate([frm_lo, [len(frm64)]]))
    surviving = n_roots - removed
    offsets = np.concatenate([[0], np.cumsum(surviving)[:-1]])
    if len(frm64) == 0:
        return frm, np.zeros(0, np.int32), offsets.astype(np.int64), frm_lo.astype(np.int64), int(surviving.sum())
    # strip-local rank of every label that occurs in a quad
    labs = np.concatenate([quads[:, 0], quads[:, 1]])
    rks = np.concatenate([quads[:, 2], quads[:, 3]]) - 1
    ulab, first = np.unique(labs, return_index=True)
    urank = rks[first]
    t_rank = urank[np.searchsorted(ulab, to64)]
    t_strip = np.searchsorted(starts, to64, side="right") - 1
    t_removed_before = np.searchsorted(frm64, to64) - frm_lo[t_strip]          # removed roots of that strip before `to`
    dense = offsets[t_strip] + t_rank - t_removed_before + 1
    return frm, dense.astype(np.int32), offsets.astype(np.int64), frm_lo.astype(np.int64), int(surviving.sum())


def first_pixel_labels(dense_labels):
    """Converts dense canonical labels (1..n in raster order of first pixel) to the sharded representation
    (1 + linear index of the region's first pixel); used to compare sharded and unsharded results."""
    flat = np.asarray(dense_labels).ravel()
    out = np.zeros_like(flat)
    pos = flat > 0
    _, first = np.unique(flat[pos], return_index=True)
    idx = np.flatnonzero(pos)[first]            # first pixel of label k+1 (labels are 1..n, sorted by np.unique)
    out[pos] = (idx + 1)[flat[pos] - 1]
    return out.reshape(np.asarray(dense_labels).shape)


def dense_from_first_pixel(labels):
    """Inverse of first_pixel_labels on a full (gathered) label image: dense 1..n by ascending first pixel."""
    flat = np.asarray(labels).ravel()
    out = np.zeros_like(flat)
    pos = flat > 0
    u, inv = np.unique(flat[pos], return_inverse=True)
    out[pos] = inv + 1
    return len(u), out.reshape(np.asarray(labels).shape)


def allgather_pairs(dist, pairs, device=None, cap=None):
    """All-gathers variable-length (n,2) int32 pair lists with torch.distributed (NCCL on GPU tensors, gloo on CPU).
    Returns the concatenation over ranks as a numpy (m,2) array, identical on every rank.
    cap: an upper bound on n known to every rank (e.g. the image width for seam pairs) -> ONE collective on a fixed-size
    buffer whose row 0 carries the count; without it the counts are gathered first (two collectives)."""
    import torch
    world = dist.get_world_size()
    arr = np.asarray(pairs, dtype=np.int32).reshape(-1, 2)
    if cap is not None:
        if len(arr) > cap:
            raise ValueError("allgather_pairs: %d pairs exceed cap %d" % (len(arr), cap))
        buf = np.zeros((cap + 1, 2), np.int32)
        buf[0, 0] = len(arr)
        buf[1:1 + len(arr)] = arr
        t = torch.from_numpy(buf)
        if device is not None:
            t = t.to(device)
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        host = torch.stack(out).cpu().numpy()
        parts = [host[r, 1:1 + int(host[r, 0, 0])] for r in range(world)]
        return np.concatenate(parts, axis=0) if parts else np.zeros((0, 2), np.int32)
    t = torch.as_tensor(arr)
    if device is not None:
        t = t.to(device)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    cap = max(1, max(counts))
    padded = torch.zeros((cap, 2), dtype=torch.int32, device=t.device)
    padded[:t.shape[0]] = t
    out = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(out, padded)
    parts = [o[:c].cpu().numpy() for o, c in zip(out, counts)]
    return np.concatenate(parts, axis=0) if parts else np.zeros((0, 2), np.int32)
