"""Synthetic RLdata-shaped records for the 1 M / 10 M benchmark configurations
(BASELINE.json configs[2:]; recipe of SURVEY.md 8d).

Shape of the reference's RLdata tables (R/data.R:1-60): five attributes - year / month / day of
birth (categorical) and first / last name (strings compared with a thresholded Levenshtein
distance, README.md:76-89) - with 10 % of the records being noisy duplicates.  Entities are
drawn i.i.d.; every record copies its entity and distorts each attribute independently; names
are distorted by one random edit, dates by a uniformly different value; a small share of the
name cells is missing.  Everything is generated as integer codes with numpy so that 10 M rows
take seconds."""
from __future__ import annotations

import numpy as np

from .model import (Attribute, BetaRV, CategoricalAttribute, Levenshtein, exchanger_coded,
                    transform_dist_fn)

_ONSETS = ["B", "D", "F", "G", "H", "K", "L", "M", "N", "R", "S", "T", "W", "SCH"]
_VOWELS = ["A", "E", "I", "O", "U"]
_CODAS = ["", "", "", "N", "R", "S"]
_LETTERS = "ABCDEFGHIJKLMNOPQRSTUVWXYZ"


def _make_names(rng, count: int) -> list[str]:
    names, seen = [], set()
    while len(names) < count:
        k = 3
        s = "".join(_ONSETS[rng.integers(len(_ONSETS))] + _VOWELS[rng.integers(len(_VOWELS))] +
                    _CODAS[rng.integers(len(_CODAS))] for _ in range(k))
        if len(s) <= 14 and s not in seen:
            seen.add(s)
            names.append(s)
    return names


def _typo(rng, s: str) -> str:
    pos = int(rng.integers(len(s)))
    kind = int(rng.integers(3))
    c = _LETTERS[rng.integers(26)]
    if kind == 0:
        return s[:pos] + c + s[pos + 1:]
    if kind == 1:
        return s[:pos] + c + s[pos:]
    return s[:pos] + s[pos + 1:] if len(s) > 2 else s[:pos] + c + s[pos + 1:]


def _name_domain(rng, base_count: int):
    """base names plus one typo variant each; returns (domain, variant_of[base] -> code)."""
    base = _make_names(rng, base_count)
    domain = list(base)
    lut = {s: i for i, s in enumerate(domain)}
    variant = np.zeros(base_count, np.int32)
    for i, s in enumerate(base):
        t = _typo(rng, s)
        if t not in lut:
            lut[t] = len(domain)
            domain.append(t)
        variant[i] = lut[t]
    return domain, variant


def synth_rldata(n_records: int, seed: int = 20260101, dup_frac: float = 0.1, distort: float = 0.03,
                 na_frac: float = 0.01, d_fname: int | None = None, d_lname: int | None = None,
                 name_skew: float = 0.5):
    """Returns ``(codes [5, N] int32, domains, attr_names, identity [N])`` with attributes ordered
    by, bm, bd, fname, lname (categorical first, as ``exchanger()`` sorts them).

    ``name_skew`` is the Zipf exponent of the name frequencies.  It sets how many DISTINCT entities
    share both names: at 9 M entities about 6.4 others per entity for 0.9, 0.04 for 0.5.  The model
    treats a shared rare full name as overwhelming evidence (dates are cheap to distort), so with
    dense name collisions its posterior merges all namesakes, theta of the dates drifts to 0.2-0.3
    over a few hundred sweeps and a sweep costs 30x more (profiles/README.md, "chain drift").
    0.5 keeps the collision density below RLdata10000's, where the reference reaches F1 0.94."""
    rng = np.random.default_rng(seed)
    N = int(n_records)
    M = max(1, int(round((1.0 - dup_frac) * N)))
    d_fname = d_fname or int(np.clip(N // 50, 200, 20000))
    d_lname = d_lname or int(np.clip(N // 20, 400, 50000))
    dom_f, var_f = _name_domain(rng, d_fname)
    dom_l, var_l = _name_domain(rng, d_lname)

    def zipf(d):
        w = 1.0 / (np.arange(d) + 10.0) ** name_skew
        return w / w.sum()

    years = np.arange(1920, 2010)
    wy = np.exp(-0.5 * ((years - 1965) / 22.0) ** 2)
    ent = np.stack([
        rng.choice(len(years), M, p=wy / wy.sum()),
        rng.integers(0, 12, M),
        rng.integers(0, 28, M),
        rng.choice(d_fname, M, p=zipf(d_fname)),
        rng.choice(d_lname, M, p=zipf(d_lname)),
    ]).astype(np.int32)
    identity = np.concatenate([np.arange(M), rng.integers(0, M, N - M)])
    identity = identity[rng.permutation(N)]
    codes = ent[:, identity].copy()
    sizes = [len(years), 12, 28]
    for a in range(3):  # categorical distortion: a uniformly different value
        hit = rng.random(N) < distort
        shift = rng.integers(1, sizes[a], N)
        codes[a] = np.where(hit, (codes[a] + shift) % sizes[a], codes[a])
    for a, var in ((3, var_f), (4, var_l)):  # names: the typo variant of the base name
        hit = rng.random(N) < distort
        codes[a] = np.where(hit, var[codes[a]], codes[a])
        codes[a] = np.where(rng.random(N) < na_frac, -1, codes[a])
    domains = [[str(y) for y in years], [str(m + 1) for m in range(12)], [str(d + 1) for d in range(28)],
               dom_f, dom_l]
    return codes.astype(np.int32), domains, ["by", "bm", "bd", "fname_c1", "lname_c1"], identity


def synth_model(n_records: int, clust_prior, seed: int = 20260101, distort_prior=None,
                neighbour_fn=None, **kw):
    """An :class:`ExchangERModel` over :func:`synth_rldata`.  ``neighbour_fn(domain, threshold)``
    computes the Levenshtein neighbourhoods (default: the GPU builder of the C ABI); the README
    attribute specification is used (threshold 3, exclude_entity_value)."""
    codes, domains, names, identity = synth_rldata(n_records, seed=seed, **kw)
    prior = distort_prior or BetaRV(1, 50)
    lev = transform_dist_fn(Levenshtein(), threshold=3)
    specs = [CategoricalAttribute(prior), CategoricalAttribute(prior), CategoricalAttribute(prior),
             Attribute(lev, prior), Attribute(lev, prior)]
    if neighbour_fn is None:
        from .capi import levenshtein_neighbours
        neighbour_fn = levenshtein_neighbours
    nbrs = {names[a]: neighbour_fn(domains[a], 3) for a in (3, 4)}
    return exchanger_coded(codes, domains, names, specs, clust_prior, neighbours=nbrs), identity
