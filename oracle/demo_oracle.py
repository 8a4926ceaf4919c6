"""TEST INFRASTRUCTURE -- CPU restatement of the notebook's single-point `enkf` (efa_demo.ipynb cell 11, lines 25-93).

Only tests/ may import this module (it is the checker of efa_xray_b200.demo.enkf); the product never does.
Pinned by construction: it is the notebook's statements in the notebook's order, with the shuffle (lines 44-46)
replaced by an explicit order so that a run can be reproduced.
"""
import numpy as np


def enkf_numpy(obs, prior, obs_range=(1, 2), ob_error=1.0, inflation=1.0, order=None):
    """efa_demo.ipynb cell 11 with the shuffle replaced by a given order."""
    prior = np.asarray(prior, dtype=np.float64)
    Nstate, Nens = prior.shape
    sel = list(obs[obs_range[0] - 1:obs_range[-1]])
    post_mean = prior.mean(axis=1)
    post_pert = (prior - post_mean[:, None]) * inflation
    for obnum in (range(len(sel)) if order is None else order):
        ob = sel[obnum]
        ob_idx = obs_range[0] + obnum - 1
        prior_mean, prior_pert = post_mean, post_pert
        H = np.zeros(Nstate)
        H[ob_idx] = 1.0
        ye = np.dot(H, prior_pert + prior_mean[:, None])
        ye_mean = np.mean(ye)
        ye_variance = np.var(ye - ye_mean)
        innov = ob - ye_mean
        kcov = np.dot(prior_pert, ye) / (Nens - 1)
        K = kcov / (ye_variance + ob_error)
        post_mean = prior_mean + K * innov
        beta = 1.0 / (1.0 + np.sqrt(ob_error / (ye_variance + ob_error)))
        post_pert = prior_pert - np.dot((beta * K)[:, None], (ye - ye_mean)[None, :])
    return post_pert + post_mean[:, None]
