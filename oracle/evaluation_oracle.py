"""CPU restatement of the reference's ranking evaluation -- TEST INFRASTRUCTURE ONLY.

Follows implicit/evaluation.pyx:366-475 (`ranking_metrics_at_k`) user by user and rank by rank in plain
Python, and :14-48 (`train_test_split`).  Pinned against the reference's own compiled module
(oracle/_ref/evaluation*.so) by tests/test_oracle.py and against tests/golden/eval_metrics.npz.
"""
import math

import numpy as np
from scipy.sparse import csr_matrix


def ranking_metrics_at_k(model, train_user_items, test_user_items, K=10, batch_size=1000):
    train_user_items = train_user_items.tocsr()
    test_user_items = test_user_items.tocsr()
    users, items = test_user_items.shape
    cg = [1.0 / math.log2(i) for i in range(2, K + 2)]  # :393
    cg_sum = np.cumsum(cg)  # :394
    relevant = pr_div = total = mean_ap = ndcg = mean_auc = 0.0
    indptr, indices = test_user_items.indptr, test_user_items.indices
    to_generate = np.arange(users, dtype="int32")[np.ediff1d(indptr) > 0]  # :421-422
    for start in range(0, len(to_generate), batch_size):  # :426-429
        batch = to_generate[start:start + batch_size]
        ids, _ = model.recommend(batch, train_user_items[batch], N=K)
        for b, u in enumerate(batch):  # :432
            likes = set(int(i) for i in indices[indptr[u]:indptr[u + 1]])  # :435-438
            pr_div += min(K, len(likes))  # :440
            ap = hit = miss = auc = 0.0
            idcg = cg_sum[min(K, len(likes)) - 1]  # :446
            num_pos = len(likes)
            num_neg = items - num_pos
            for i in range(K):  # :450-458
                if int(ids[b, i]) in likes:
                    relevant += 1
                    hit += 1
                    ap += hit / (i + 1)
                    ndcg += cg[i] / idcg
                else:
                    miss += 1
                    auc += hit
            auc += ((hit + num_pos) / 2.0) * (num_neg - miss)  # :459
            mean_ap += ap / min(K, len(likes))  # :460
            mean_auc += auc / (num_pos * num_neg)  # :461
            total += 1
    return {"precision": relevant / pr_div, "map": mean_ap / total, "ndcg": ndcg / total,
            "auc": mean_auc / total}  # :469-474


def train_test_split(ratings, train_percentage=0.8, random_state=None):
    ratings = ratings.tocoo()  # :31
    rng = np.random.default_rng(random_state)  # :32 (check_random_state on an int / None)
    random_index = rng.random(len(ratings.data))  # :33
    out = []
    for sel in (random_index < train_percentage, random_index >= train_percentage):  # :34-43
        out.append(csr_matrix((ratings.data[sel], (ratings.row[sel], ratings.col[sel])), shape=ratings.shape,
                              dtype=ratings.dtype))
    train, test = out
    test.data[test.data < 0] = 0  # :45-46
    test.eliminate_zeros()
    return train, test
