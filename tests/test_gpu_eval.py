"""GPU parity of the widened rows: evaluation steps, the periodic report, weight checkpoints
and the main.py-compatible CLI, all through the C-ABI.

  test_step_lg_vae / test_step_lg_gm_vae   vae/trainer.py:199-274   (forward + loss terms, no update)
  model.save_weights                        vae/trainer.py:421
"""
import numpy as np
import pytest
import torch

from oracle import splitvae_oracle as O
from helpers import make_case, rel_l2, to_dev

pytestmark = pytest.mark.gpu


def _model(kind, H, precision="fp32"):
    from splitvae_b200.model import LGGMVae, LGVae
    if kind == "lgvae":
        return LGVae(128, 128, image_shape=[-1, H, H, 3], precision=precision)
    return LGGMVae(128, 128, [-1, H, H, 3], 30, 0.4, precision=precision)


@pytest.mark.parametrize("kind", ["lgvae", "lggmvae"])
def test_eval_step_matches_oracle_and_leaves_weights_alone(kind):
    from splitvae_b200 import trainer
    H, B, beta, alpha = 32, 4, 40.0, 40.0
    params, batch = make_case(kind, H, B, 4, seed_base=40)
    m = _model(kind, H)
    m.configure(beta=beta, alpha=alpha)
    m.set_weights_by_name(params)
    m.build(B)
    before = m.engine.params.clone()
    x, eg, el = to_dev(batch["inputs"]), to_dev(batch["eps_g"]), to_dev(batch["eps_l"])
    cfg = {"beta": beta, "alpha": alpha}
    metrics = trainer.make_metrics()
    if kind == "lgvae":
        sc = trainer.test_step_lg_vae(m, x, config=cfg, metrics=metrics, eps_g=eg, eps_l=el)
    else:
        outs = trainer.test_step_lg_gm_vae(m, x, config=cfg, metrics=metrics, eps_g=eg, eps_l=el, u=to_dev(batch["u"]))
        assert len(outs) == 14                                     # the reference's return tuple (trainer.py:274)
        sc = trainer.test_step_lg_gm_vae.last_scalars
    ref_sc, _ = O.forward_backward(params, kind, batch["inputs"], batch["eps_g"], batch["eps_l"],
                                   batch["u"] if kind == "lggmvae" else None, beta=beta, alpha=alpha, dtype=torch.float64)
    for k, v in ref_sc.items():
        assert abs(sc[k] - v) <= 1e-5 * max(1.0, abs(v)), (k, sc[k], v)
    assert torch.equal(before, m.engine.params)                     # evaluation never touches the weights
    assert m.engine.iterations == 0
    assert abs(metrics["x_recon_test_loss"].result() - ref_sc["recon_x"]) <= 1e-5 * abs(ref_sc["recon_x"])
    text = trainer.format_report(0, metrics)
    assert text.startswith("Training step 0") and "Test X Recon Loss" in text


def test_eval_engine_for_other_batch_size_tracks_training_weights():
    """The evaluation batch may differ from the training batch: a second engine is fed the training engine's weights."""
    from splitvae_b200 import trainer
    kind, H = "lgvae", 32
    params, batch = make_case(kind, H, 4, 4, seed_base=41)
    m = _model(kind, H)
    m.configure(beta=1.0)
    m.set_weights_by_name(params)
    m.build(2)                                                       # "training" engine: batch 2
    m.engine.params.mul_(1.01)                                       # pretend a train step moved the weights
    m.engine.params_updated()
    x, eg, el = to_dev(batch["inputs"]), to_dev(batch["eps_g"]), to_dev(batch["eps_l"])
    sc = trainer.test_step_lg_vae(m, x, config={"beta": 1.0}, eps_g=eg, eps_l=el)          # evaluation batch 4
    moved = {k: v * np.float32(1.01) for k, v in params.items()}
    ref_sc, _ = O.forward_backward(moved, kind, batch["inputs"], batch["eps_g"], batch["eps_l"], None, beta=1.0, dtype=torch.float64)
    assert abs(sc["total"] - ref_sc["total"]) <= 2e-5 * abs(ref_sc["total"])
    assert m.engine.B == 2


@pytest.mark.parametrize("kind", ["lgvae", "lggmvae"])
def test_checkpoint_round_trip_resumes_bit_exactly(kind, tmp_path):
    """save_weights(include_optimizer) -> fresh model -> load_weights: the next train step is bit-identical."""
    H, B = 32, 4
    params, batch = make_case(kind, H, B, 4, seed_base=42)
    x, eg, el = to_dev(batch["inputs"]), to_dev(batch["eps_g"]), to_dev(batch["eps_l"])
    u = to_dev(batch["u"]) if kind == "lggmvae" else None
    a = _model(kind, H)
    a.set_weights_by_name(params)
    a.build(B)
    for _ in range(2):
        a.engine.train_step(x, eg, el, u)
    path = a.save_weights(str(tmp_path / "ckpt"), include_optimizer=True)
    with np.load(path) as z:
        for name, shape, _, _ in a.engine.table:                    # Keras names and layouts
            assert z[name].shape == tuple(shape)
        assert int(z["optimizer/iterations"]) == 2
    a.engine.train_step(x, eg, el, u)
    torch.cuda.synchronize()
    b = _model(kind, H)
    b.build(B)
    b.load_weights(path)
    assert b.engine.iterations == 2
    b.engine.train_step(x, eg, el, u)
    torch.cuda.synchronize()
    assert torch.equal(a.engine.params, b.engine.params)
    assert torch.equal(a.engine.adam_v, b.engine.adam_v)


def test_cli_trains_evaluates_and_saves(tmp_path, capsys):
    """`main.py --model lgvae --dataset svhn ...` with the reference's flags: train loop, evaluation report, weights file."""
    from splitvae_b200 import main as cli
    out = tmp_path / "w"
    hist = cli.main(["--model", "lgvae", "--dataset", "svhn", "--beta", "1", "--patch_size", "4", "--batch_size", "8", "-no_label",
                     "--training_steps", "6", "--report_every", "3", "--test_batches", "2", "--precision", "fp32",
                     "--save_weights", str(out)])
    text = capsys.readouterr().out
    assert text.count("Training step") == 3 and "Testing time" in text and "Training done!" in text
    assert len(hist) == 3 and all(np.isfinite(sc["total"]) for _, sc in hist)
    assert hist[-1][1]["total"] < hist[0][1]["total"]                # six Adam steps on a fixed synthetic pool reduce the loss
    with np.load(str(out) + ".npz") as z:
        assert "encoder_x.e1.kernel" in z.files or any(k.endswith("e1.kernel") for k in z.files)
