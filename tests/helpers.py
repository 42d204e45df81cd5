"""Seeded synthetic inputs shared by the oracle and the GPU parity tests."""
from typing import List, NamedTuple, Tuple

import numpy as np
import scipy.sparse as sps


class FMWeights(NamedTuple):
    global_bias: float
    weight: np.ndarray
    factors: np.ndarray


STUB_WEIGHT = FMWeights(
    -3.0,
    np.asarray([1.0, 2.0, -1.0]),
    np.asarray([[1.0, -1.0, 0], [0.0, 1.0, 1.0], [1.0, 1.0, 1.0], [-1.0, 0, -1.0]]),
)


def fm_prediction(X: sps.csr_matrix, weight: FMWeights) -> np.ndarray:
    """Plain numpy FM forward pass (reference tests/test_utils.py:15-25)."""
    X2 = X.copy()
    X2.data[:] = X2.data ** 2
    result = np.full(X.shape[0], weight.global_bias, dtype=np.float64)
    result += X.dot(weight.weight)
    w2 = (weight.factors ** 2).sum(axis=0)
    Xw = X.dot(weight.factors.T)
    result += ((Xw ** 2).sum(axis=1) - X2.dot(w2)) * 0.5
    return result


def middle_data(n_train: int = 1000) -> Tuple[sps.csr_matrix, np.ndarray]:
    """The planted 3-feature fixture of the reference (tests/conftest.py:26-45)."""
    rns = np.random.RandomState(0)
    rows: List[int] = []
    cols: List[int] = []
    data: List[float] = []
    for row in range(n_train):
        for ind in np.where(rns.random(3) > 0.5)[0]:
            rows.append(row)
            cols.append(ind)
            data.append(float(rns.choice([-2, -1, 1, 2])))
    X = sps.csr_matrix((data, (rows, cols)), shape=(n_train, 3))
    return X, fm_prediction(X, STUB_WEIGHT)


def toy_matrix() -> Tuple[sps.csr_matrix, np.ndarray]:
    """The literal 4 x 9 matrix of the reference README (README.md:56-59) — config C1."""
    X = np.asarray(
        [
            [19.0, 0, 0, 0, 1, 1, 0, 0, 0],
            [33.0, 0, 0, 1, 0, 0, 1, 0, 0],
            [55.0, 0, 1, 0, 0, 0, 0, 1, 0],
            [20.0, 1, 0, 0, 0, 0, 0, 0, 1],
        ]
    )
    return sps.csr_matrix(X), np.asarray([0.0, 1.0, 1.0, 0.0])


def movielens_like(n_rows: int, n_users: int, n_movies: int, rank: int, seed: int, noise: float = 0.9,
                   zipf=0.8):
    """Two one-hot fields per row (user, movie) with power-law popularity and a planted
    rank-`rank` FM + Gaussian noise: the shape of the MovieLens configs.
    `zipf` is the popularity exponent, one value or (users, movies): 0.8 is a heavy tail (the
    most popular movie gets several per cent of all rows); (0.4, 0.45) reproduces the max/mean
    ratings-per-user (51x) and per-movie (37x) ratios of the public ML-10M statistics."""
    rng = np.random.default_rng(seed)
    zu, zm = (zipf, zipf) if np.isscalar(zipf) else zipf

    def popularity(n, s):
        p = 1.0 / np.arange(1, n + 1) ** s
        return p / p.sum()

    users = rng.choice(n_users, size=n_rows, p=popularity(n_users, zu))
    movies = rng.choice(n_movies, size=n_rows, p=popularity(n_movies, zm))
    # every user / movie appears at least once when there is room
    if n_rows >= n_users + n_movies:
        users[:n_users] = rng.permutation(n_users)
        movies[n_users:n_users + n_movies] = rng.permutation(n_movies)
    indptr = np.arange(0, 2 * n_rows + 1, 2, dtype=np.int64)
    indices = np.empty(2 * n_rows, dtype=np.int32)
    indices[0::2] = users
    indices[1::2] = n_users + movies
    X = sps.csr_matrix((np.ones(2 * n_rows), indices, indptr), shape=(n_rows, n_users + n_movies))
    bu, bm = rng.normal(0, 0.3, n_users), rng.normal(0, 0.3, n_movies)
    Vu, Vm = rng.normal(0, 0.5, (n_users, rank)), rng.normal(0, 0.5, (n_movies, rank))
    y = 3.5 + bu[users] + bm[movies] + (Vu[users] * Vm[movies]).sum(1) + rng.normal(0, noise, n_rows)
    return X, y, [n_users, n_movies]


def block_data(n_train: int = 100, seed: int = 0):
    """The relation-block fixture of the reference (tests/regression/test_block.py:80-118):
    returns (X_flatten, tm_column, (user_indices, user_block), (item_indices, item_block), y,
    group_shapes)."""
    rns = np.random.RandomState(seed)
    user_block = sps.csr_matrix(np.eye(3), dtype=np.float64)
    user_indices = rns.randint(0, user_block.shape[0], size=n_train)
    item_block = sps.csr_matrix(np.eye(2), dtype=np.float64)
    group_shapes = [1, user_block.shape[1], item_block.shape[1]]
    item_indices = rns.randint(0, item_block.shape[0], size=n_train)
    tm_column = rns.randn(n_train, 1)
    X_flatten = sps.hstack([tm_column, user_block[user_indices], item_block[item_indices]]).tocsr()
    X2 = X_flatten.copy()
    X2.data = X2.data ** 2
    weights = rns.randn(3, X_flatten.shape[1])
    Xw = X_flatten.dot(weights.T)
    y = 0.5 * ((Xw ** 2).sum(axis=1) - X2.dot((weights ** 2).sum(axis=0))) + rns.randn(n_train)
    return (X_flatten, sps.csr_matrix(tm_column), (user_indices, user_block),
            (item_indices, item_block), y, group_shapes)


def dense_block_data(n_train: int = 400, seed: int = 3):
    """Relation blocks with shared (non one-hot) columns, so block columns conflict through block
    rows and the block level schedule has several levels (SVD++-style implicit features)."""
    rns = np.random.RandomState(seed)
    n_user, n_item = 12, 9
    user_implicit = (rns.random((n_user, n_item)) < 0.4).astype(np.float64)
    user_implicit /= np.maximum(1.0, np.sqrt(user_implicit.sum(1, keepdims=True)))
    user_block = sps.hstack([sps.eye(n_user), sps.csr_matrix(user_implicit)]).tocsr()
    item_side = rns.randn(n_item, 2)
    item_block = sps.hstack([sps.eye(n_item), sps.csr_matrix(item_side)]).tocsr()
    user_indices = rns.randint(0, n_user, size=n_train)
    item_indices = rns.randint(0, n_item, size=n_train)
    main = sps.csr_matrix((rns.random((n_train, 2)) < 0.5) * rns.randn(n_train, 2))
    group_shapes = [2, n_user, n_item, n_item, 2]
    X_flatten = sps.hstack([main, user_block[user_indices], item_block[item_indices]]).tocsr()
    w = rns.randn(X_flatten.shape[1]) * 0.3
    F = rns.randn(2, X_flatten.shape[1]) * 0.4
    X2 = X_flatten.copy()
    X2.data = X2.data ** 2
    y = (X_flatten.dot(w) + 0.5 * (((X_flatten.dot(F.T)) ** 2).sum(1) - X2.dot((F ** 2).sum(0)))
         + 0.5 * rns.randn(n_train))
    return X_flatten, main, (user_indices, user_block), (item_indices, item_block), y, group_shapes


def fields_like(n_rows: int, field_sizes, rank: int, seed: int, unit: bool = True, zipf: float = 0.8,
                noise: float = 0.5):
    """L categorical fields per row (one active category each, sorted column order), power-law
    popularity; values 1 (`unit`) or random in [0.5, 1.5].  Planted rank-`rank` FM + noise."""
    rng = np.random.default_rng(seed)
    L = len(field_sizes)
    offsets = np.concatenate([[0], np.cumsum(field_sizes)])
    cols = np.empty((n_rows, L), dtype=np.int32)
    for f, n in enumerate(field_sizes):
        p = 1.0 / np.arange(1, n + 1) ** zipf
        c = rng.choice(n, size=n_rows, p=p / p.sum())
        if n_rows >= n:
            where = rng.permutation(n_rows)[:n]
            c[where] = rng.permutation(n)  # every category appears
        cols[:, f] = offsets[f] + c
    vals = np.ones((n_rows, L)) if unit else rng.uniform(0.5, 1.5, size=(n_rows, L))
    indptr = np.arange(0, L * n_rows + 1, L, dtype=np.int64)
    X = sps.csr_matrix((vals.ravel(), cols.ravel(), indptr), shape=(n_rows, int(offsets[-1])))
    w = rng.normal(0, 0.3, X.shape[1])
    F = rng.normal(0, 0.4, (rank, X.shape[1]))
    X2 = X.copy()
    X2.data = X2.data ** 2
    y = 1.0 + X.dot(w) + 0.5 * ((X.dot(F.T) ** 2).sum(1) - X2.dot((F ** 2).sum(0))) + rng.normal(0, noise, n_rows)
    return X, y, list(field_sizes)


def ml1m_extended(n_rows, n_users, n_movies, n_days, seed):
    """Main table: day one-hot; user block: id one-hot + implicit feedback (movies rated, 1/sqrt n);
    movie block: id one-hot + implicit (users who rated it)."""
    rng = np.random.default_rng(seed)
    users = rng.integers(0, n_users, n_rows)
    movies = rng.integers(0, n_movies, n_rows)
    days = rng.integers(0, n_days, n_rows)
    R = sps.csr_matrix((np.ones(n_rows), (users, movies)), shape=(n_users, n_movies))
    R.data[:] = 1.0
    Ru = sps.diags(1.0 / np.sqrt(np.maximum(1, np.asarray(R.sum(1)).ravel()))) @ R
    Rm = sps.diags(1.0 / np.sqrt(np.maximum(1, np.asarray(R.sum(0)).ravel()))) @ R.T
    user_block = sps.hstack([sps.eye(n_users), Ru]).tocsr()
    movie_block = sps.hstack([sps.eye(n_movies), Rm]).tocsr()
    main = sps.csr_matrix((np.ones(n_rows), (np.arange(n_rows), days)), shape=(n_rows, n_days))
    y = 3.5 + rng.normal(0, 0.3, n_users)[users] + rng.normal(0, 0.3, n_movies)[movies] + rng.normal(0, 0.9, n_rows)
    return main, (users, user_block), (movies, movie_block), y, [n_days, n_users, n_movies, n_movies, n_users]
