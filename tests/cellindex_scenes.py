"""Labelled-cell scenes for the S2CellIndex tests (modelled on s2cell_index_test.cc) — shared by the CPU and GPU tests."""
import numpy as np


def debug_id(s):
    """S2CellId::FromDebugString("f/ddd") (s2cell_id.cc:600-640)."""
    face, digits = s.split("/")
    id_ = (int(face) << 61) | (1 << 60)
    for d in digits:
        new_lsb = (id_ & -id_) >> 2
        id_ = id_ + (2 * int(d) + 1 - 4) * new_lsb
    return id_


def parent(ids, level):
    ids = np.asarray(ids, np.uint64)
    lsb = np.uint64(1) << (np.uint64(2) * (np.uint64(30) - np.asarray(level, np.uint64)))
    return (ids & (~lsb + np.uint64(1))) | lsb


def random_leaves(rng, n):
    """uniform leaf cell ids (face, 60 random position bits)"""
    face = rng.integers(0, 6, n).astype(np.uint64)
    pos = rng.integers(0, 1 << 60, n, dtype=np.uint64)
    return (face << np.uint64(61)) | (pos << np.uint64(1)) | np.uint64(1)


def random_cells(rng, n, min_level=0, max_level=30):
    return parent(random_leaves(rng, n), rng.integers(min_level, max_level + 1, n))


def normalize_sorted_disjoint(ids):
    """sorted, duplicates and contained cells dropped (what S2CellUnion requires of a target; siblings are not merged)"""
    ids = np.unique(np.asarray(ids, np.uint64))
    lsb = ids & (~ids + np.uint64(1))
    lo, hi = ids - (lsb - np.uint64(1)), ids + (lsb - np.uint64(1))
    order = np.lexsort((np.iinfo(np.uint64).max - hi, lo))     # by range_min, larger cells first
    out, cur_hi = [], None
    for k in order:
        if cur_hi is not None and hi[k] <= cur_hi:
            continue
        out.append(ids[k]); cur_hi = hi[k]
    return np.array(sorted(out), np.uint64)


def random_union(rng, n_cells):
    return normalize_sorted_disjoint(random_cells(rng, n_cells))


def clustered_scene(rng, n_pairs, n_labels, n_seeds=40):
    """cells at all levels concentrated around a few seed leaves, so nesting, duplicates and shared range_min are common"""
    seeds = random_leaves(rng, n_seeds)
    base = seeds[rng.integers(0, n_seeds, n_pairs)]
    # perturb the low bits of some, keep others identical (nested chains along one leaf)
    jitter = rng.integers(0, 1 << 20, n_pairs, dtype=np.uint64) << np.uint64(1)
    keep = rng.integers(0, 3, n_pairs) == 0
    leaves = np.where(keep, base, (base & ~np.uint64((1 << 21) - 1)) | jitter | np.uint64(1))
    ids = parent(leaves, rng.integers(0, 31, n_pairs))
    labels = rng.integers(0, n_labels, n_pairs).astype(np.int32)
    return ids, labels


def semi_random_scene(rng):
    """IntersectionSemiRandomUnions (s2cell_index_test.cc:416-434): a walk along the curve that adds / targets cells."""
    id_ = debug_id("1/0123012301230123")
    ids, labels, target = [], [], []
    for i in range(100):
        if rng.random() < 0.1:
            ids.append(id_); labels.append(i)
        if rng.random() < 0.25:
            target.append(id_)
        lsb = id_ & -id_
        if rng.random() < 0.5:
            id_ = id_ + (lsb << 1)                      # next_wrap
            if id_ >= (6 << 61):
                id_ -= (6 << 61)
        if rng.random() < 1 / 6 and lsb < (1 << 60):    # parent
            nl = lsb << 2
            id_ = (id_ & -nl) | nl
            lsb = nl
        if rng.random() < 1 / 6 and lsb > 1:            # child_begin
            id_ = id_ - lsb + (lsb >> 2)
    return np.array(ids, np.uint64), np.array(labels, np.int32), normalize_sorted_disjoint(np.array(target, np.uint64))


def sort_pairs_per_target(off, cells, labels):
    """canonical order within each target: by (cell id, label)"""
    cells, labels = cells.copy(), labels.copy()
    for t in range(len(off) - 1):
        b, e = int(off[t]), int(off[t + 1])
        o = np.lexsort((labels[b:e], cells[b:e]))
        cells[b:e] = cells[b:e][o]; labels[b:e] = labels[b:e][o]
    return cells, labels
