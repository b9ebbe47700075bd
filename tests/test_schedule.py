"""Host logic: schedule tables, respacing and coefficient table vs the reference's own numbers
(tests/golden/schedule.npz was produced by running the reference, oracle/make_golden.py)."""
import hashlib
import os

import numpy as np
import pytest
import torch

from audio2photoreal_b200.schedule import (DiffusionTables, named_beta_schedule, respaced_betas, space_timesteps,
                                           step_coefficients)
from audio2photoreal_b200.sampler import Sampler, create_gaussian_diffusion
from argparse import Namespace

TABLES = ["betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
          "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2"]


def _sampler(resp):
    return create_gaussian_diffusion(Namespace(noise_schedule="cosine", timestep_respacing=resp, sigma_small=True,
                                               data_format="pose", model_path="x"))


@pytest.mark.parametrize("tag,resp", [("full", ""), ("ddim500", "ddim500"), ("ddim100", "ddim100"), ("ddim10", "ddim10"),
                                      ("sec10", "10"), ("sec25_10", "25,10")])
def test_tables_bit_identical_to_reference(golden_dir, tag, resp):
    g = np.load(os.path.join(golden_dir, "schedule.npz"))
    s = _sampler(resp)
    assert list(s.timestep_map) == list(g[f"{tag}/timestep_map"])
    for t in TABLES:
        assert np.array_equal(getattr(s, t), g[f"{tag}/{t}"]), (tag, t)


def test_known_answers_survey_appendix_c():
    s = _sampler("")
    assert s.num_timesteps == 1000
    assert s.betas[0] == 4.128422482196914e-05 and s.betas[999] == 0.999
    assert s.alphas_cumprod[500] == 0.49228517244880304
    assert s.sqrt_recip_alphas_cumprod[999] == 20291.169634661146
    assert s.posterior_mean_coef2[500] == 0.9953562794552251
    assert hashlib.sha1(s.alphas_cumprod.tobytes()).hexdigest()[:12] == "e2ffecd3dbdd"
    d10 = _sampler("ddim10")
    assert d10.timestep_map == list(range(0, 1000, 100))
    assert d10.alphas_cumprod[-1] == 0.02361610879439386


def test_space_timesteps_rules_and_errors():
    assert space_timesteps(1000, "ddim500") == set(range(0, 1000, 2))
    assert sorted(space_timesteps(1000, "10"))[:3] == [0, 111, 222]
    assert space_timesteps(1000, [1000]) == set(range(1000))
    with pytest.raises(ValueError):
        space_timesteps(1000, "ddim999")    # no integer stride gives 999 steps
    with pytest.raises(ValueError):
        space_timesteps(10, "20")           # section smaller than the count


def test_step_coefficients_match_reference_ops():
    s = _sampler("ddim10")
    co = step_coefficients(s, eta=0.0)
    assert co.shape == (10, 8) and co.dtype == np.float32
    f32 = lambda a: torch.from_numpy(a).float()
    assert np.array_equal(co[:, 0], f32(s.sqrt_recip_alphas_cumprod).numpy())
    assert np.array_equal(co[:, 2], torch.sqrt(f32(s.alphas_cumprod_prev)).numpy())   # cast first, sqrt in fp32
    assert (co[:, 4] == 0).all()                        # eta = 0
    assert co[0, 7] == 0 and (co[1:, 7] > 0).all()      # no noise at t == 0
    co5 = step_coefficients(s, eta=0.5)
    assert co5[0, 4] == 0 and (co5[1:, 4] > 0).all()


def test_sampler_rejects_unsupported_configs():
    betas = named_beta_schedule("cosine", 1000)
    with pytest.raises(NotImplementedError):
        Sampler(range(1000), betas=betas, rescale_timesteps=True)
    with pytest.raises(NotImplementedError):
        named_beta_schedule("sqrt", 10)
    s = _sampler("ddim10")
    with pytest.raises(NotImplementedError):
        s.ddim_sample_loop(None, (1, 104, 1, 8), dump_steps=[1])
    with pytest.raises(NotImplementedError):
        s.ddim_sample_loop(None, (1, 104, 1, 8), const_noise=True)
