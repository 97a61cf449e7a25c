#!/usr/bin/env python
"""Synthetic books.csv / users.csv in the reference's schema for BASELINE configs[0].

    python scripts/make_config1_data.py --out /tmp/cfg1 --books 2000 --users 600 --per-user 24

The shipped data/*_trimmed.csv share no ASINs (last row), so configs[0] runs on generated files:
  books.csv : title,author,average_rating,rating_number,price,categories,parent_asin
              categories = '["Books", "<c1>", "<c2>"]' from a 40 x 8 vocabulary, ~300 authors
  users.csv : parent_asin,userId,timestamp   (Int64 milliseconds)
Users have two preferred top-level categories and draw 80 % of their books from them (so that retrieval metrics are not
flat); every user has >= 3 interactions and popular books clear the reference's min_item_interactions = 6 filter
(default.yaml:16-17).  Deterministic for a given seed.
"""
from __future__ import annotations

import argparse
import json
from pathlib import Path

import numpy as np
import pandas as pd


def generate(out: Path, *, books: int = 2000, users: int = 600, per_user: int = 24, seed: int = 7,
             top: int = 40, sub: int = 8, authors: int = 300) -> dict:
    rng = np.random.default_rng(seed)
    out = Path(out)
    out.mkdir(parents=True, exist_ok=True)
    c1 = rng.integers(0, top, books)
    c2 = rng.integers(0, sub, books)
    auth = np.minimum((rng.pareto(1.2, books) * 8).astype(np.int64), authors - 1)
    asin = np.array([f"B{seed:02d}{i:07d}" for i in range(books)])
    cats = [json.dumps(["Books", f"Genre {a:02d}", f"Genre {a:02d} / Shelf {b}"]) for a, b in zip(c1, c2)]
    frame = pd.DataFrame({
        "title": [json.dumps([f"Volume {i} of genre {a}"]) for i, a in enumerate(c1)],
        "author": [f"Author {a:03d}" for a in auth],
        "average_rating": np.round(rng.uniform(2.5, 5.0, books), 1),
        "rating_number": rng.integers(1, 5000, books),
        "price": np.round(rng.gamma(2.0, 6.0, books), 2),
        "categories": cats,
        "parent_asin": asin,
    })
    frame.to_csv(out / "books.csv", index=False)
    by_cat = [np.flatnonzero(c1 == a) for a in range(top)]
    pop = 1.0 / np.arange(1, books + 1) ** 0.6
    pop = pop[rng.permutation(books)]
    rows_a, rows_u, rows_t = [], [], []
    t0 = 1_600_000_000_000
    for u in range(users):
        n = max(3, int(rng.poisson(per_user)))
        fav = rng.choice(top, size=2, replace=False)
        pool = np.concatenate([by_cat[f] for f in fav])
        k_fav = min(len(pool), int(round(0.8 * n)))
        w = pop[pool] / pop[pool].sum()
        picks = set(rng.choice(pool, size=k_fav, replace=False, p=w).tolist()) if k_fav > 0 else set()
        while len(picks) < n:
            picks.add(int(rng.choice(books, p=pop / pop.sum())))
        picks = sorted(picks)
        ts = np.sort(rng.integers(t0, t0 + 50_000_000_000, len(picks)))
        uid = f"U{seed:02d}{u:08d}"
        rows_a += asin[picks].tolist()
        rows_u += [uid] * len(picks)
        rows_t += ts.tolist()
    inter = pd.DataFrame({"parent_asin": rows_a, "userId": rows_u, "timestamp": pd.array(rows_t, dtype="Int64")})
    inter = inter.sample(frac=1.0, random_state=seed).reset_index(drop=True)
    inter.to_csv(out / "users.csv", index=False)
    return {"books": books, "users": users, "interactions": len(inter)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--books", type=int, default=2000)
    ap.add_argument("--users", type=int, default=600)
    ap.add_argument("--per-user", type=int, default=24)
    ap.add_argument("--seed", type=int, default=7)
    a = ap.parse_args()
    print(json.dumps(generate(Path(a.out), books=a.books, users=a.users, per_user=a.per_user, seed=a.seed)))
