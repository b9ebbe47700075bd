"""N2 / N3: the KV-cached guide sampler and the VQ decoder against the reference's own modules
(model/guide.py GuideTransformer, model/vqvae.py TemporalVertexCodec) on CPU, with reference-layout checkpoints.
Host-side PyTorch: no GPU needed.  Skipped when no reference view is importable."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ref_harness as RH

pytestmark = pytest.mark.skipif(not RH.reference_available(), reason="no reference view (python -m oracle.build_ref)")


def _ref_guide(tokens=32, layers=2, dim=64):
    ref = RH.import_reference()
    import model.guide as G
    torch.manual_seed(5)
    with RH._cwd(ref.scratch):
        m = G.GuideTransformer(tokens=tokens, num_layers=layers, dim=dim, emb_len=798, num_audio_layers=2).eval()
    for p in m.parameters():                       # non-trivial biases / affine so every term is exercised
        if p.dim() == 1:
            torch.nn.init.normal_(p, 0.0 if "bias" in "" else 0.0, 0.05)
    return m


def _audio(B, frames=240, seed=3):
    g = torch.Generator().manual_seed(seed)
    return 0.1 * torch.randn(B, frames * 1600, 2, generator=g)


def test_guide_sampler_logits_and_generate_match_reference():
    from audio2photoreal_b200.guide import GuideSampler
    m = _ref_guide()
    ours = GuideSampler(m.state_dict(), tokens=m.tokens, audio_model=m.audio_model).eval()
    assert set(ours.state_dict()) == set(m.state_dict())          # same checkpoint layout, frozen extractor included
    B, n = 2, 12
    cond = _audio(B)
    g = torch.Generator().manual_seed(9)
    toks = torch.cat([torch.full((B, 1), m.tokens), torch.randint(0, m.tokens, (B, n - 1), generator=g)], dim=1)
    with torch.no_grad():
        ref_logits = m(toks, cond)
    got = ours.logits_for(toks, cond)
    assert torch.allclose(got, ref_logits, rtol=1e-4, atol=2e-5), (got - ref_logits).abs().max().item()

    # generate(): identical token sequences under the same uniform tape (inverse-CDF draw on both sides)
    tape = torch.rand(64, B, generator=torch.Generator().manual_seed(11))

    def make_draw():
        it = iter(tape)
        return lambda probs: (torch.cumsum(probs, -1) < next(it).unsqueeze(-1)).sum(-1).clamp(max=probs.shape[-1] - 1)
    from torch.distributions import Categorical
    d_ref = make_draw()
    old = Categorical.sample
    Categorical.sample = lambda self, *a, **k: d_ref(self.probs)
    try:
        ref_tok = m.generate(cond, sequence_length=4, layers=3, n_sequences=B)
    finally:
        Categorical.sample = old
    got_tok = ours.generate(cond, sequence_length=4, layers=3, n_sequences=B, draw=make_draw())
    assert got_tok.shape == ref_tok.shape == (B, 12) and torch.equal(got_tok, ref_tok)


def test_vq_decoder_matches_reference_and_checkpoint_layout(tmp_path):
    from audio2photoreal_b200.guide import VQDecoder, setup_tokenizer
    RH.import_reference()
    import model.vqvae as V
    torch.manual_seed(2)
    codec = V.TemporalVertexCodec(n_vertices=104, latent_dim=64, categories=32, residual_depth=4).eval()
    for layer in codec.quantizer.layers:
        layer._codebook.embed.normal_()            # kmeans_init leaves zeros until training
    d = tmp_path / "vq"
    d.mkdir()
    with open(d / "args.json", "w") as f:
        json.dump({"nb_joints": 104, "output_emb_width": 64, "code_dim": 32, "depth": 4}, f)
    torch.save({"net": codec.state_dict()}, d / "net_iter.pth")
    ours = setup_tokenizer(str(d / "net_iter.pth"), device="cpu")
    assert isinstance(ours, VQDecoder) and ours.residual_depth == 4 and ours.n_clusters == 32
    q = torch.randint(0, 32, (3, 20, 4))
    with torch.no_grad():
        want = codec.decode(q)
    got = ours.decode(q)
    assert got.shape == want.shape == (3, 20, 104)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)


def test_results_block_layout(tmp_path):
    """sample/generate.py:146-152,289-292: np.save of a dict with these five keys, re-loadable with allow_pickle"""
    from audio2photoreal_b200.guide import inv_transform, results_block, save_results
    stats = {"pose_mean": np.full(104, 0.1, np.float32), "pose_std_flat": np.float32(0.5), "code_mean": np.zeros(256, np.float32),
             "code_std_flat": np.float32(2.0), "audio_mean": np.zeros(2, np.float32), "audio_std_flat": np.float32(3.0)}
    s = torch.randn(2, 104, 1, 60)
    motion = inv_transform(s.permute(0, 2, 3, 1), "pose", stats).permute(0, 3, 1, 2)
    assert torch.allclose(motion, s * 0.5 + 0.1)
    blk = results_block([motion], [np.zeros((2, 96000, 2), np.float32)], [motion], [torch.full((2,), 60)], [torch.zeros(2, 2, 104)])
    save_results(str(tmp_path / "out" / "results.npy"), blk)
    back = np.load(tmp_path / "out" / "results.npy", allow_pickle=True).item()
    assert set(back) == {"motions", "audio", "gt", "lengths", "keyframes"} and back["motions"].shape == (2, 104, 1, 60)
