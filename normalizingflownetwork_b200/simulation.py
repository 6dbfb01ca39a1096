"""Synthetic conditional-density data for demos, tests and the config-1 benchmark.

Same distributions as the reference's simulation/dummy_data_gen.py (cosine + heteroscedastic
noise, Trippe-style heteroscedastic sine), written against numpy.random.Generator-free legacy
seeding so that seed 22 reproduces the reference's draws."""
import numpy as np


def gen_cosine_noise_data(num, noise_std=0.2, heterosced_noise=0.0):
    """x on [-4, 4]; y = 4 sin(x) + N(0, noise_std) + |x| * N(0, heterosced_noise) on the first half."""
    np.random.seed(22)
    x = np.linspace(-4, 4, num=num)
    hetero = np.where(np.arange(num) < int(num / 2), np.abs(x), 0.0)
    base_noise = np.random.normal(0, noise_std, size=num)
    het_noise = np.random.normal(0, heterosced_noise, size=num)
    y = 4 * np.sin(x) + base_noise + hetero * het_noise
    return x.astype(np.float32).reshape((num, 1)), y.astype(np.float32).reshape((num, 1))


def gen_trippe_hetero_data(dim=1, n_pts=10000, bimodal=False, heteroscedastic=True, asymetric=False):
    """1-D output, heteroscedastic (optionally bimodal / asymmetric) sine data on [-pi, pi]^dim."""
    np.random.seed(22)
    lo, hi = -np.pi, np.pi
    noise_scale = 3.0 if heteroscedastic else 0.0
    n_mode = int(n_pts / 2.0) if bimodal else n_pts

    def draw(fold):
        X = np.random.uniform(lo, hi, size=[n_mode, dim])
        std = noise_scale * (np.abs(np.sin(X)).prod(axis=1) if fold else np.abs(np.sin(X).prod(axis=1)))
        eps = np.random.normal(0.0, np.abs(std))
        mean = 5.0 * (np.sin(np.abs(X).prod(axis=1)) if fold else np.sin(X).prod(axis=1))
        return X, (np.abs(eps) if asymetric else eps) + mean

    X, Y = draw(False)
    if bimodal:
        X2, Y2 = draw(True)
        X, Y = np.concatenate([X, X2]), np.concatenate([Y, Y2])
    return X, Y.reshape([n_pts, 1])
