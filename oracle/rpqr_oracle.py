"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's accelerated RPQR (rpqr.py:7-67) with explicit random
streams, used as the checker of randomly_pivoted_cholesky_b200/rpqr.py.  Pinned to the live reference by
tests/golden/rpqr_*.npz (tests/golden/make_golden_rpqr.py).  Never imported by the product package."""
import numpy as np

from .rpc_oracle import rejection_cholesky, sample_pivots


def accelerated_rpqr(B, k, b, rng, legacy):
    """rpqr.py:7-67 with a FIXED b; `rng` = proposal Generator, `legacy` = RandomState standing in for np.random.rand()."""
    m, n = B.shape
    Q = np.zeros((m, k))
    F = np.zeros((n, k))
    u = np.linalg.norm(B, axis=0) ** 2                                        # :16
    arr_idx = np.zeros(k, dtype=int)
    counter = 0
    rounds = []
    while counter < k:
        idx = sample_pivots(u, rng.random(b))                                 # :27
        C = B[:, idx] - Q[:, :counter] @ F[idx, :counter].T                   # :32
        _, accepted = rejection_cholesky(C.T @ C, legacy.random_sample(b))    # :33
        num_sel = len(accepted)
        if num_sel > k - counter:                                             # :41-43
            num_sel = k - counter
            accepted = accepted[:num_sel]
        idx = idx[accepted]
        C = C[:, accepted]
        new_counter = counter + len(accepted)
        arr_idx[counter:new_counter] = idx
        Q[:, counter:new_counter], _ = np.linalg.qr(C - Q[:, :counter] @ (Q[:, :counter].T @ C), mode="reduced")  # :49
        F[:, counter:new_counter] = B.T @ Q[:, counter:new_counter]           # :50
        u -= np.linalg.norm(F[:, counter:new_counter], axis=1) ** 2           # :59
        u = u.clip(min=0)                                                     # :60
        counter = new_counter
        rounds.append(num_sel)
    return Q, F, arr_idx, rounds
