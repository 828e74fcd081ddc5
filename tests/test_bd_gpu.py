"""Path C parity on the GPU: gc_bd_posterior_f32/f64 through the C-ABI against the reference
dumps (1e-5 absolute, BASELINE.json north_star) and against the oracle on random batches."""
import os

import numpy as np
import pytest
import torch

import gym_cooking_b200 as gcb
import oracle as O

pytestmark = pytest.mark.gpu


def _run(g, dtype):
    dev = "cuda"
    probs = torch.tensor(g["prior"], dtype=dtype, device=dev)
    out = gcb.bd_posterior(probs, torch.tensor(g["alive"], device=dev), torch.tensor(g["hyp_pair"], device=dev),
                           torch.tensor(g["pair_w"], device=dev), torch.tensor(g["qdiff"], dtype=dtype, device=dev),
                           torch.tensor(g["n_valid"], device=dev), torch.tensor(g["act_idx"], device=dev),
                           float(g["beta"]))
    return out.cpu().numpy().astype(np.float64)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.float64, 1e-12)])
def test_reference_posteriors(golden_dir, dtype, tol):
    f = np.load(os.path.join(golden_dir, "bd_posteriors.npz"))
    g = {k: f[k] for k in f.files}
    post = _run(g, dtype)
    assert np.abs(post - g["posterior"]).max() < tol


@pytest.mark.parametrize("H,P,A,E", [(8, 8, 5, 2), (3, 4, 5, 2), (16, 12, 25, 3), (36, 15, 25, 3), (84, 24, 25, 4),
                                     (96, 128, 32, 4), (1, 1, 1, 1)])
def test_random_batches_match_oracle(H, P, A, E):
    rng = np.random.RandomState(H * 1000 + P)
    n = 3001
    g = dict(prior=rng.rand(n, H), alive=(rng.rand(n, H) < 0.85).astype(np.uint8),
             hyp_pair=rng.randint(0, P, size=(n, H, E)).astype(np.uint8), pair_w=rng.randint(1, 3, size=(n, P)).astype(np.uint8),
             qdiff=rng.randn(n, P, A) * 2, n_valid=rng.randint(1, A + 1, size=(n, P)).astype(np.uint8), beta=1.3)
    g["hyp_pair"][rng.rand(n, H, E) < 0.3] = 255
    g["act_idx"] = (rng.randint(0, 1 << 16, size=(n, P)) % g["n_valid"]).astype(np.uint8)
    g["alive"][0] = 0  # a row with no surviving hypothesis
    g["prior"][1] = 0  # a row whose total is zero -> uniform
    expect = O.bd_posterior(g["prior"], g["alive"], g["hyp_pair"], g["pair_w"], g["qdiff"], g["n_valid"],
                            g["act_idx"], 1.3)
    assert np.abs(_run(g, torch.float64) - expect).max() < 1e-12
    assert np.abs(_run(g, torch.float32) - expect).max() < 1e-5
