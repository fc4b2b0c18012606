"""Adversarial inputs for the Sign / CrossingSign chain: configurations that force the stable and exact stages."""
import numpy as np


def _norm(p):
    return p / np.sqrt((p * p).sum(axis=-1))[..., None]


def sign_triples(n, seed):
    """(m, 9) float64 rows a|b|c mixing: random, nearly collinear, exactly collinear (z = 0 plane and axis planes),
    proportional points, tiny clusters, tiny/huge magnitudes, duplicated points."""
    rng = np.random.default_rng(seed)
    out = []
    a = _norm(rng.standard_normal((n, 3))); b = _norm(rng.standard_normal((n, 3))); c = _norm(rng.standard_normal((n, 3)))
    out.append(np.hstack([a, b, c]))
    # c nearly on the great circle through a, b
    t = rng.uniform(-1, 2, (n, 1))
    c2 = _norm(a * (1 - t) + b * t)
    out.append(np.hstack([a, b, c2]))
    # tiny clusters
    eps = 10.0 ** rng.uniform(-15, -6, (n, 1))
    out.append(np.hstack([a, _norm(a + eps * rng.standard_normal((n, 3))), _norm(a + eps * rng.standard_normal((n, 3)))]))
    # exactly collinear: all in the plane z = 0 (determinant exactly zero)
    ang = rng.uniform(0, 2 * np.pi, (n, 3))
    pts = [np.stack([np.cos(ang[:, k]), np.sin(ang[:, k]), np.zeros(n)], axis=1) for k in range(3)]
    out.append(np.hstack(pts))
    # same with x = 0 / y = 0 and small-integer coordinates (many exact zeros among minors)
    ints = rng.integers(-3, 4, (n, 9)).astype(float)
    out.append(ints)
    ints2 = ints.copy(); ints2[:, [0, 3, 6]] = 0
    out.append(ints2)
    ints3 = ints.copy(); ints3[:, [2, 5, 8]] = 0; ints3[:, 6:9] = 0
    out.append(ints3)
    # proportional points
    k = rng.choice([1 - 2 ** -53, 1 + 2 ** -52, 0.5, 2.0, -1.0, -(1 - 2 ** -53)], (n, 1))
    out.append(np.hstack([a, a * k, b]))
    out.append(np.hstack([a, b, a * k]))
    # duplicates
    out.append(np.hstack([a, a, b])); out.append(np.hstack([a, b, b])); out.append(np.hstack([a, b, a]))
    # wild magnitudes (underflow territory)
    sc = 10.0 ** rng.uniform(-320, 0, (n, 9))
    out.append(rng.choice([-1.0, 1.0, 0.0], (n, 9)) * sc)
    sc2 = 10.0 ** rng.uniform(-100, 100, (n, 9))
    out.append(rng.standard_normal((n, 9)) * sc2)
    return np.ascontiguousarray(np.vstack(out))


def crossing_quads(n, seed):
    """(m, 12) rows a|b|c|d: general position, four points on one great circle, all on the equator, shared vertices,
    degenerate edges, near-midpoint configurations."""
    rng = np.random.default_rng(seed)
    out = []
    a, b, c, d = (_norm(rng.standard_normal((n, 3))) for _ in range(4))
    out.append(np.hstack([a, b, c, d]))
    # short edges near each other (crossings are frequent)
    e = 0.05
    b2 = _norm(a + e * rng.standard_normal((n, 3))); c2 = _norm(a + e * rng.standard_normal((n, 3))); d2 = _norm(a + e * rng.standard_normal((n, 3)))
    out.append(np.hstack([a, b2, c2, d2]))
    # all four on the great circle through a, b (nearly)
    t = rng.uniform(-1, 2, (n, 2))
    out.append(np.hstack([a, b2, _norm(a * (1 - t[:, :1]) + b2 * t[:, :1]), _norm(a * (1 - t[:, 1:]) + b2 * t[:, 1:])]))
    # exactly on the equator
    ang = rng.uniform(0, 2 * np.pi, (n, 4))
    out.append(np.hstack([np.stack([np.cos(ang[:, k]), np.sin(ang[:, k]), np.zeros(n)], axis=1) for k in range(4)]))
    # shared vertices in all four positions
    out.append(np.hstack([a, b2, a, d2])); out.append(np.hstack([a, b2, c2, a]))
    out.append(np.hstack([a, b2, b2, d2])); out.append(np.hstack([a, b2, c2, b2]))
    out.append(np.hstack([a, b2, a, b2])); out.append(np.hstack([a, b2, b2, a]))
    # degenerate edges
    out.append(np.hstack([a, a, c2, d2])); out.append(np.hstack([a, b2, c2, c2])); out.append(np.hstack([a, a, a, a]))
    # c is (almost) the midpoint of ab
    out.append(np.hstack([a, b2, _norm(a + b2), d2]))
    # shared vertex plus collinear third point on the equator
    out.append(np.hstack([np.stack([np.cos(ang[:, 0]), np.sin(ang[:, 0]), np.zeros(n)], axis=1),
                          np.stack([np.cos(ang[:, 1]), np.sin(ang[:, 1]), np.zeros(n)], axis=1),
                          np.stack([np.cos(ang[:, 0]), np.sin(ang[:, 0]), np.zeros(n)], axis=1),
                          np.stack([np.cos(ang[:, 2]), np.sin(ang[:, 2]), np.zeros(n)], axis=1)]))
    return np.ascontiguousarray(np.vstack(out))
