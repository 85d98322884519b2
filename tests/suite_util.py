"""Shared knobs for the py3 restatement of the reference's test-suite
(reference: gp/tests/util.py:4-52 -- same distributions, seed and tolerances)."""
import numpy as np

OPT = dict(n_big=100, n_small=10, pct_fail=5, rtol=1e-5, dtheta=1e-5)
DTHETA = OPT["dtheta"]


def seed():
    np.random.seed(2348)                                  # util.py:47-48


def rand_params(*names):
    draw = {"h": lambda: np.random.uniform(0, 2),          # util.py:15-28
            "w": lambda: np.random.uniform(np.pi / 32., np.pi / 2.),
            "p": lambda: np.random.uniform(0.33, 3),
            "s": lambda: np.random.uniform(0, 0.5)}
    return tuple(draw[n]() for n in names)


def central(y0, y1, dx):
    return (y1 - y0) / 2. / dx                            # util.py:31-33


def make_xy():
    x = np.linspace(-2 * np.pi, 2 * np.pi, 16)            # util.py:36-39
    return x, np.sin(x)


def make_xo():
    return np.linspace(-2 * np.pi, 2 * np.pi, 32)         # util.py:42-44


def allclose(a, b):
    return np.allclose(a, b, rtol=OPT["rtol"])            # util.py:51-52


# the reference suite's one hard-coded known-answer case (gp/tests/test_gp.py:298-333):
# Kxx is numerically singular, Cholesky must fail.
INVALID_H = 0.53356762
INVALID_W = 2.14797803
INVALID_X = np.array([
    0.0, 0.3490658503988659, 0.6981317007977318, 1.0471975511965976, 1.3962634015954636,
    1.7453292519943295, 2.0943951023931953, 0.41968261, 0.97349106, 1.51630532, 1.77356282,
    2.07011378, 2.87018553, 3.70955074, 3.96680824, 4.50962249, 4.80617345, 5.06343095,
    5.6062452])
INVALID_Y = np.array([
    -5.297411814764175e-16, 2.2887507861169e-16, 1.1824308893126911e-15,
    1.9743321560961036e-15, 3.387047586844716e-15, 3.2612801348363973e-15,
    2.248201624865942e-15, -3.061735126365188e-05, 2.1539042816804896e-05,
    -3.900581031467468e-05, 4.603140942399664e-05, 0.00014852070373963522,
    -0.011659908151004955, -0.001060998167383152, -0.0002808538329216448,
    -8.057870658869265e-06, -7.668984947838558e-07, -7.910215881378919e-08,
    -3.2649468298271893e-10])
