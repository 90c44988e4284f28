"""GPU: the drop-in model classes (splitvae_b200.model.LGVae / LGGMVae) against vectors produced by
the reference's OWN vae/model.py, imported unmodified under a stand-in for tensorflow (scripts/make_reference_model_golden.py):
output-tuple ORDER of `model(inputs)`, `encode`, `decode(rescale=True/False)`, `encode_y`, `get_y`, through the C-ABI
(sv_forward, sv_decode, sv_encode_y), fp32 reference-kernel mode."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import splitvae_oracle as O
from helpers import to_dev

pytestmark = pytest.mark.gpu


def _close(t, d, what):
    t = t.detach().double().cpu().reshape(-1)
    assert t.numel() == d["n"], (what, t.numel(), d["n"])
    scale = max(1.0, d["l2"] / max(1, d["n"]) ** 0.5)
    assert abs(float(t.norm()) - d["l2"]) <= 2e-4 * max(1.0, d["l2"]), (what, float(t.norm()), d["l2"])
    assert abs(float(t.sum()) - d["sum"]) <= 2e-4 * max(1.0, abs(d["sum"]), d["l2"] * d["n"] ** 0.5), (what, float(t.sum()), d["sum"])
    assert np.allclose(t[:4].numpy(), d["head"], rtol=5e-4, atol=2e-4 * scale), (what, t[:4].tolist(), d["head"])


@pytest.mark.parametrize("kind", ["lgvae", "lggmvae"])
def test_model_classes_match_the_reference_model(kind):
    from splitvae_b200.model import LGGMVae, LGVae
    with open(os.path.join(os.path.dirname(__file__), "golden", f"reference_model_{kind}.json")) as f:
        G = json.load(f)
    c = G["case"]
    H, B = c["H"], c["B"]
    params = O.init_params(kind, H, H, seed=5 + c["seed_base"])
    b = O.synthetic_batch(B, H, c["patch"], seed_base=c["seed_base"])
    if kind == "lgvae":
        m = LGVae(128, 128, image_shape=[-1, H, H, 3], precision="fp32")
    else:
        m = LGGMVae(128, 128, [-1, H, H, 3], 30, 0.4, precision="fp32")
    assert (m.global_latent_dims, m.local_latent_dims, m.image_shape) == (128, 128, [-1, H, H, 3])
    m.set_weights_by_name(params)
    x, eg, el = to_dev(b["inputs"]), to_dev(b["eps_g"]), to_dev(b["eps_l"])
    u = to_dev(b["u"]) if kind == "lggmvae" else None
    # model(inputs): the reference's tuple, element by element IN ORDER (vae/model.py:200, 248)
    tup = m(x, eps_g=eg, eps_l=el) if kind == "lgvae" else m(x, training=True, eps_g=eg, eps_l=el, u=u)
    torch.cuda.synchronize()
    assert len(tup) == len(G["output_order"]) == (10 if kind == "lgvae" else 14)
    for t, name in zip(tup, G["output_order"]):
        _close(t, G["outputs"][name], f"call[{name}]")
    # encode -> (z_x, z_x_hat), sampled (model.py:204-209, 252-257)
    enc = m.encode(x, eg, el, u) if kind == "lggmvae" else m.encode(x, eg, el)
    for t, d, n in zip(enc, G["api"]["encode"], ("z_x", "z_x_hat")):
        _close(t.clone(), d, f"encode[{n}]")
    # decode(z_x, z_x_hat, rescale) (model.py:211-218, 259-266)
    za, zb = 0.5 * eg, 0.5 * el
    for t, d, n in zip(m.decode(za, zb), G["api"]["decode_rescaled"], ("x_recon", "x_hat_recon")):
        _close(t, d, f"decode(rescale=True)[{n}]")
        assert float(t.min()) >= 0.0 and float(t.max()) <= 1.0
    for t, d, n in zip(m.decode(za, zb, rescale=False), G["api"]["decode_raw"], ("x_mean", "x_hat_mean")):
        _close(t, d, f"decode(rescale=False)[{n}]")
    if kind == "lggmvae":
        y_in = torch.softmax(torch.log(u.double()), dim=1).float().contiguous()
        for t, d, n in zip(m.encode_y(y_in), G["api"]["encode_y"], ("z_prior_mean", "z_prior_sig")):
            _close(t, d, f"encode_y[{n}]")
        m.engine.forward(x, eg, el, u)          # get_y draws its own noise in the reference; here the same u is injected
        for t, d, n in zip((m.engine.output("y"), m.engine.output("y_logits")), G["api"]["get_y"], ("y", "y_logits")):
            _close(t, d, f"get_y[{n}]")
        y, y_logits = m.get_y(x, u=u)
        _close(y, G["api"]["get_y"][0], "get_y()[y]")
        _close(y_logits, G["api"]["get_y"][1], "get_y()[y_logits]")
        assert m.y_size == 30


def test_gmvae_class_matches_the_reference_model():
    """--model gmvae (vae/model.py:277-320): 9-tuple order of `model(inputs)`, encode -> z_x, decode(z_x, rescale), encode_y, get_y,
    against the vectors of the reference's own GMVae class (tests/golden/reference_model_gmvae.json)."""
    from splitvae_b200.model import GMVae
    with open(os.path.join(os.path.dirname(__file__), "golden", "reference_model_gmvae.json")) as f:
        G = json.load(f)
    c = G["case"]
    H, B = c["H"], c["B"]
    params = O.init_params("gmvae", H, H, seed=5 + c["seed_base"])
    b = O.synthetic_batch(B, H, c["patch"], seed_base=c["seed_base"])
    m = GMVae(128, [-1, H, H, 3], 30, 0.4, precision="fp32")
    assert (m.global_latent_dims, m.image_shape, m.y_size) == (128, [-1, H, H, 3], 30)
    m.set_weights_by_name(params)
    x, eg, u = to_dev(b["inputs"]), to_dev(b["eps_g"]), to_dev(b["u"])
    tup = m(x, training=True, eps_g=eg, u=u)
    torch.cuda.synchronize()
    assert len(tup) == len(G["output_order"]) == 9
    for t, name in zip(tup, G["output_order"]):
        _close(t, G["outputs"][name], f"call[{name}]")
    api = G.get("api", {})
    if "encode" in api:
        _close(m.encode(x, eg, u).clone(), api["encode"][0] if isinstance(api["encode"], list) else api["encode"], "encode[z_x]")
    z = 0.5 * eg
    rec = m.decode(z)
    assert tuple(rec.shape) == (B, H, H, 3) and float(rec.min()) >= 0.0 and float(rec.max()) <= 1.0
    raw = m.decode(z, rescale=False).clone()
    assert torch.allclose(torch.clip((raw + 1) * 0.5, 0., 1.), m.decode(z), atol=1e-6)
    P = O.to_torch(params, torch.float64, requires_grad=False)       # decode / encode_y against the oracle's restatement
    ref_mean, _ = O.decoder(P, "decoder_x", z.double().cpu(), H, H)
    assert float((raw.double().cpu() - ref_mean).norm() / ref_mean.norm()) < 1e-4
    y_in = torch.softmax(torch.log(u.double()), dim=1).float().contiguous()
    zpm, zps = m.encode_y(y_in)
    yd = y_in.double().cpu()
    ref_pm = yd @ P["encoder_x.z_prior_mean.kernel"] + P["encoder_x.z_prior_mean.bias"]
    ref_ps = torch.nn.functional.softplus(yd @ P["encoder_x.z_prior_sig.kernel"] + P["encoder_x.z_prior_sig.bias"])
    assert float((zpm.double().cpu() - ref_pm).norm() / ref_pm.norm()) < 1e-4 and float((zps.double().cpu() - ref_ps).norm() / ref_ps.norm()) < 1e-4
    y, y_logits = m.get_y(x, u=u)
    _close(y, G["outputs"]["y"], "get_y()[y]")
    _close(y_logits, G["outputs"]["y_logits"], "get_y()[y_logits]")
