"""Pins the CPU oracle for path C against posteriors dumped from the reference's
BayesianDelegator.bayes_update (oracle/gen_golden.py bd): tolerance 1e-12 for the float64
oracle (the north star allows 1e-5 for the product)."""
import os

import numpy as np

import oracle as O


def load(golden_dir):
    g = np.load(os.path.join(golden_dir, "bd_posteriors.npz"))
    return {k: g[k] for k in g.files}


def test_posterior_oracle_matches_reference(golden_dir):
    g = load(golden_dir)
    post = O.bd_posterior(g["prior"], g["alive"], g["hyp_pair"], g["pair_w"], g["qdiff"], g["n_valid"],
                          g["act_idx"], float(g["beta"]))
    assert post.shape[0] >= 50
    assert np.abs(post - g["posterior"]).max() < 1e-12
    alive = g["alive"].astype(bool)
    assert np.allclose(post.sum(axis=1), 1.0) and (post[~alive] == 0).all()


def test_known_answer_from_survey():
    """SURVEY.md section 8a worked example: None-likelihood softmax(1.3*[0.5, 1/6, 1/6, 1/6])[0] = 0.3396."""
    qd = np.zeros((1, 1, 4))
    qd[0, 0] = [0.5, 1 / 6, 1 / 6, 1 / 6]
    post = O.bd_posterior(np.array([[0.25, 0.75]]), None, np.array([[[0, 255]], [[255, 255]]]).reshape(1, 2, 2),
                          np.array([[1]]), qd, np.array([[4]]), np.array([[0]]), 1.3)
    L = np.exp(0.65) / (np.exp(0.65) + 3 * np.exp(1.3 / 6))
    assert abs(L - 0.3396) < 1e-4
    assert np.allclose(post, [[1.0, 0.0]])  # second hypothesis has no entries -> factor 0


def test_zero_total_falls_back_to_uniform():
    post = O.bd_posterior(np.array([[0.0, 0.0, 0.0]]), np.array([[1, 0, 1]]), np.full((1, 3, 1), 0),
                          np.array([[1]]), np.zeros((1, 1, 2)), np.array([[2]]), np.array([[0]]), 1.3)
    assert np.allclose(post, [[0.5, 0.0, 0.5]])
