"""GPU parity on randomly drawn problems (seeded): graph shape, degree skew, D, dtype, ops, weights, layout path.
Every draw is checked forward + backward against the C oracle with the north-star tolerance."""
import numpy as np
import pytest
import torch

from helpers import OPS, oracle_all
from test_rspmm_gpu import run_ours, check_against

pytestmark = pytest.mark.gpu


def draw_case(rng):
    N = int(rng.choice([1, 2, 7, 33, 130, 700, 950, 2500]))
    R = int(rng.choice([1, 2, 5, 40, 300, 850]))
    dens = float(rng.choice([0.0, 0.5, 3.0, 20.0, 80.0]))
    E = int(min(N * dens, 60000))
    dtype = torch.float64 if rng.random() < 0.25 else torch.float32
    D = int(rng.choice([1, 3, 4, 16, 32, 36, 64, 100, 160, 256]))
    skew = float(rng.choice([0.0, 1.0, 2.0]))
    g = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
    if skew > 0 and E > 0:
        p = 1.0 / torch.arange(1, N + 1, dtype=torch.float64).pow(skew)
        rows = torch.multinomial(p, E, replacement=True, generator=g)
        cols = torch.multinomial(p, E, replacement=True, generator=g)[torch.randperm(E, generator=g)]
    else:
        rows = torch.randint(0, N, (E,), generator=g)
        cols = torch.randint(0, N, (E,), generator=g)
    ei = torch.stack([rows, cols])
    et = torch.randint(0, R, (E,), generator=g)
    unit = rng.random() < 0.4
    w = torch.ones(E, dtype=dtype) if unit else torch.rand(E, generator=g, dtype=dtype) + 0.5
    rel = torch.randn(R, D, generator=g, dtype=dtype)
    x = torch.randn(N, D, generator=g, dtype=dtype)
    og = torch.randn(N, D, generator=g, dtype=dtype)
    return ei, et, w, rel, x, og


@pytest.mark.parametrize("seed", range(40))
def test_random_problem(cuda_device, seed):
    rng = np.random.default_rng(1000 + seed)
    ei, et, w, rel, x, og = draw_case(rng)
    sum, mul = OPS[int(rng.integers(len(OPS)))]
    ours = run_ours(sum, mul, ei, et, w, rel, x, og)
    check_against(sum, mul, ours, oracle_all(sum, mul, ei, et, w, rel, x, og),
                  "seed %d N=%d E=%d R=%d D=%d %s %s-%s " % (seed, x.shape[0], ei.shape[1], rel.shape[0], x.shape[1],
                                                              x.dtype, sum, mul))
