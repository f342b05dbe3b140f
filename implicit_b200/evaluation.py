"""Ranking evaluation: the main batch consumer of ``recommend``.

Mirrors ``implicit.evaluation`` (implicit/evaluation.pyx): ``ranking_metrics_at_k`` :366-475 and its four
single-metric wrappers :236-363, plus ``train_test_split`` :14-48.  The reference walks the users in
batches of 1000 through ``model.recommend`` and scores each returned id against an ``unordered_set`` of
the user's withheld items in a scalar loop.  Here the batches are sized for the fused top-k kernel (one
launch covers ``BATCH`` users) and the per-user bookkeeping is whole-batch array arithmetic on the
returned ``(users, K)`` id matrix; the four sums are the same quantities, accumulated in float64.

``leave_k_out_split`` (:141-233) is dataset preparation rather than a caller of the hot path; it is provided with the
reference's contract (every eligible user leaves exactly K entries in the test matrix, train + test == input, a
reserved share of users stays train-only, the same ValueErrors).  The reference draws its shuffle from numpy's
global generator there (:128), so the split itself is seeded here and is not bit-comparable.
"""
import numpy as np
from scipy.sparse import csr_matrix

from .utils import check_random_state

#: users per ``recommend`` call (the reference uses 1000, evaluation.pyx:417; any value gives the same sums)
BATCH = 16384


def train_test_split(ratings, train_percentage=0.8, random_state=None):
    """Random split of the stored entries into (train, test) CSR matrices (evaluation.pyx:14-48).

    One uniform draw per stored entry, in COO order: ``< train_percentage`` goes to train.  Negative
    test entries are dropped (:44-45)."""
    ratings = ratings.tocoo()
    rng = check_random_state(random_state)
    draw = rng.random(len(ratings.data))
    in_train = draw < train_percentage
    parts = []
    for keep in (in_train, ~in_train):
        parts.append(csr_matrix((ratings.data[keep], (ratings.row[keep], ratings.col[keep])),
                                shape=ratings.shape, dtype=ratings.dtype))
    train, test = parts
    test.data[test.data < 0] = 0
    test.eliminate_zeros()
    return train, test


def leave_k_out_split(ratings, K=1, train_only_size=0.0, random_state=None):
    """'Leave K out': every eligible user (more than K + 1 stored entries, evaluation.pyx:191) gives K randomly
    chosen entries to the test matrix; `train_only_size` is the fraction of users kept out of the test set
    (:194-198).  Returns (train, test) CSR matrices of the input's shape with train + test == ratings."""
    if K < 1:
        raise ValueError("The 'K' must be >= 1.")
    if not 0.0 <= train_only_size < 1.0:
        raise ValueError("The 'train_only_size' must be in the range (0.0 <= x < 1.0).")
    ratings = ratings.tocoo()
    rng = check_random_state(random_state)
    users, items, data = ratings.row, ratings.col, ratings.data
    unique_users, counts = np.unique(users, return_counts=True)
    candidate = counts > K + 1
    if train_only_size > 0.0:
        size = max(1, int(len(unique_users) * train_only_size))  # _choose, :73-75
        reserved = rng.choice(len(unique_users), size=size, replace=False)
        # positions (not ids) drawn from range(len(unique_users)) are matched against the ids, as in :195-197
        candidate &= ~np.isin(unique_users, reserved)
    in_candidates = np.isin(users, unique_users[candidate])
    # K random entries per candidate user: shuffle the entries, stable-sort by user, take the first K of each run
    pos = np.flatnonzero(in_candidates)
    pos = pos[rng.permutation(len(pos))]
    pos = pos[np.argsort(users[pos], kind="stable")]
    run_start = np.flatnonzero(np.r_[True, users[pos][1:] != users[pos][:-1]]) if len(pos) else np.zeros(0, dtype=int)
    rank_in_run = np.arange(len(pos)) - np.repeat(run_start, np.diff(np.r_[run_start, len(pos)]))
    test_pos = pos[rank_in_run < K]
    is_test = np.zeros(len(users), dtype=bool)
    is_test[test_pos] = True
    test = csr_matrix((data[is_test], (users[is_test], items[is_test])), shape=ratings.shape, dtype=ratings.dtype)
    train = csr_matrix((data[~is_test], (users[~is_test], items[~is_test])), shape=ratings.shape, dtype=ratings.dtype)
    return train, test


def _withheld_keys(test_user_items):
    """Sorted unique ``user * items + item`` keys of the test matrix and the per-user count of DISTINCT
    withheld items (the reference's ``likes`` is a set, evaluation.pyx:436-438: duplicates count once)."""
    users, items = test_user_items.shape
    indptr = np.asarray(test_user_items.indptr, dtype=np.int64)
    rows = np.repeat(np.arange(users, dtype=np.int64), np.diff(indptr))
    keys = np.unique(rows * items + np.asarray(test_user_items.indices, dtype=np.int64))
    distinct = np.bincount(keys // items, minlength=users) if len(keys) else np.zeros(users, dtype=np.int64)
    return keys, distinct


def ranking_metrics_at_k(model, train_user_items, test_user_items, K=10, show_progress=True, num_threads=1):
    """precision / map / ndcg / auc at K over the users that have withheld items (evaluation.pyx:366-475).

    ``model.recommend(batch, train_user_items[batch], N=K)`` supplies the ranked ids (liked training
    items filtered, the recommend default).  Per user, with ``L`` distinct withheld items and
    ``hit_i`` = "rank i is withheld":
      precision  sum hit_i / sum min(K, L)                                     (:440, :451-452, :470)
      map        mean over users of  sum_i hit_i * (hits so far) / (i+1) / min(K, L)          (:453, :460)
      ndcg       mean of  sum_i hit_i / log2(i+2) / sum_{j<min(K,L)} 1/log2(j+2)              (:454, :446)
      auc        mean of [sum_{miss i} hits before i + (hits + L)/2 * (items - L - misses)] / (L (items - L))
                                                                                (:455-461)
    """
    if not isinstance(train_user_items, csr_matrix):
        train_user_items = train_user_items.tocsr()
    if not isinstance(test_user_items, csr_matrix):
        test_user_items = test_user_items.tocsr()
    K = int(K)
    users, items = test_user_items.shape
    keys, distinct = _withheld_keys(test_user_items)
    to_generate = np.arange(users, dtype=np.int32)[np.ediff1d(test_user_items.indptr) > 0]  # :421-422

    gain = 1.0 / np.log2(np.arange(2, K + 2))
    gain_sum = np.cumsum(gain)
    rank = np.arange(1, K + 1, dtype=np.float64)

    relevant = pr_div = mean_ap = ndcg = mean_auc = 0.0
    total = 0
    progress = None
    if show_progress:
        from tqdm.auto import tqdm

        progress = tqdm(total=len(to_generate))
    for start in range(0, len(to_generate), BATCH):
        batch = to_generate[start:start + BATCH]
        ids, _ = model.recommend(batch, train_user_items[batch], N=K)
        ids = np.asarray(ids).reshape(len(batch), -1).astype(np.int64)
        width = ids.shape[1]  # == K unless the model returned fewer columns
        q = batch.astype(np.int64)[:, None] * items + ids
        pos = np.searchsorted(keys, q)
        hit = np.zeros(q.shape, dtype=bool)
        inside = pos < len(keys)
        hit[inside] = keys[pos[inside]] == q[inside]
        hitf = hit.astype(np.float64)
        L = distinct[batch].astype(np.float64)
        cap = np.minimum(float(K), L)

        cum_hits = np.cumsum(hitf, axis=1)
        hits = cum_hits[:, -1] if width else np.zeros(len(batch))
        misses = float(K) - hits  # ranks the model did not fill count as misses, as in the scalar loop
        relevant += hits.sum()
        pr_div += cap.sum()
        mean_ap += ((hitf * cum_hits / rank[:width]).sum(axis=1) / cap).sum()
        ndcg += ((hitf * gain[:width]).sum(axis=1) / gain_sum[cap.astype(np.int64) - 1]).sum()
        neg = items - L
        # a miss at rank i adds the hits seen so far; unfilled ranks (width < K) all come after every hit
        auc = ((1.0 - hitf) * cum_hits).sum(axis=1) + (K - width) * hits
        auc += (hits + L) / 2.0 * (neg - misses)
        mean_auc += (auc / (L * neg)).sum()
        total += len(batch)
        if progress is not None:
            progress.update(len(batch))
    if progress is not None:
        progress.close()
    return {
        "precision": relevant / pr_div,
        "map": mean_ap / total,
        "ndcg": ndcg / total,
        "auc": mean_auc / total,
    }


def precision_at_k(model, train_user_items, test_user_items, K=10, show_progress=True, num_threads=1):
    """evaluation.pyx:236-267"""
    return ranking_metrics_at_k(model, train_user_items, test_user_items, K, show_progress, num_threads)["precision"]


def mean_average_precision_at_k(model, train_user_items, test_user_items, K=10, show_progress=True, num_threads=1):
    """evaluation.pyx:270-299"""
    return ranking_metrics_at_k(model, train_user_items, test_user_items, K, show_progress, num_threads)["map"]


def ndcg_at_k(model, train_user_items, test_user_items, K=10, show_progress=True, num_threads=1):
    """evaluation.pyx:302-331"""
    return ranking_metrics_at_k(model, train_user_items, test_user_items, K, show_progress, num_threads)["ndcg"]


def AUC_at_k(model, train_user_items, test_user_items, K=10, show_progress=True, num_threads=1):
    """evaluation.pyx:334-363"""
    return ranking_metrics_at_k(model, train_user_items, test_user_items, K, show_progress, num_threads)["auc"]
