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
    from splitvae_b200.model import GMVae, LGGMVae, LGVae
    if kind == "lgvae":
        return LGVae(128, 128, image_shape=[-1, H, H, 3], precision=precision)
    if kind == "gmvae":
        return GMVae(128, [-1, H, H, 3], 30, 0.4, precision=precision)
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


@pytest.mark.parametrize("kind", ["lgvae", "lggmvae", "gmvae"])
def test_keras_hdf5_checkpoint_round_trip(kind, tmp_path):
    """vae/trainer.py:421 `model.save_weights('models/<run>.h5')`: a path ending in .h5 writes the Keras HDF5 layout (hdf5_lite); a fresh
    model built from it resumes bit-exactly (weights, Adam moments, iteration count), and a weights-only file loads into a new model."""
    from splitvae_b200 import hdf5_lite
    from splitvae_b200.model import keras_weight_names
    H, B = 32, 4
    params, batch = make_case(kind, H, B, 4, seed_base=44)
    x, eg = to_dev(batch["inputs"]), to_dev(batch["eps_g"])
    el = to_dev(batch["eps_l"]) if kind != "gmvae" else None
    u = to_dev(batch["u"]) if kind != "lgvae" else None
    a = _model(kind, H)
    a.set_weights_by_name(params)
    a.build(B)
    for _ in range(2):
        a.engine.train_step(x, eg, el, u)
    path = a.save_weights(str(tmp_path / "run.h5"), include_optimizer=True)
    plain = a.save_weights(str(tmp_path / "weights_only.h5"))
    flat, f = hdf5_lite.load_keras_weights(plain)
    names = keras_weight_names(kind)
    assert "optimizer_weights" not in f.keys() and len(flat) == len(names) == len(a.engine.table)
    now = a.engine.get_params()
    for name, shape, _, _ in a.engine.table:                        # Keras layouts (conv HWIO, dense [in,out], bias [out]) under Keras names
        assert flat[names[name]].shape == tuple(shape) and np.array_equal(flat[names[name]], now[name])
    a.engine.train_step(x, eg, el, u)
    torch.cuda.synchronize()
    b = _model(kind, H)
    b.load_weights(path)                                            # no engine yet: applied by build()
    b.build(B)
    assert b.engine.iterations == 2
    b.engine.train_step(x, eg, el, u)
    torch.cuda.synchronize()
    assert torch.equal(a.engine.params, b.engine.params)
    assert torch.equal(a.engine.adam_m, b.engine.adam_m) and torch.equal(a.engine.adam_v, b.engine.adam_v)
    c = _model(kind, H)
    c.load_weights(plain)
    c.build(B)
    assert c.engine.iterations == 0 and all(np.array_equal(v, now[k]) for k, v in c.engine.get_params().items())


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


def test_cli_writes_the_reference_h5_and_resumes_from_it(tmp_path, capsys):
    """--save_weights <run>.h5 = vae/trainer.py:421's file; --load_weights continues from it (step count and Adam state included)."""
    from splitvae_b200 import hdf5_lite
    from splitvae_b200 import main as cli
    out = str(tmp_path / "run.h5")
    common = ["--model", "lggmvae", "--dataset", "svhn", "--beta", "40", "--alpha", "40", "--patch_size", "4", "--batch_size", "8",
              "--report_every", "2", "--test_batches", "0", "--precision", "fp32"]
    cli.main(common + ["--training_steps", "4", "--save_weights", out])
    flat, f = hdf5_lite.load_keras_weights(out)
    assert [n.decode() for n in f.attrs["layer_names"]] == ["encoder", "encoder_1", "decoder", "decoder_1"]
    assert "lggm_vae/encoder/y_dense/kernel:0" in flat and flat["lggm_vae/decoder_1/conv2d_13/kernel:0"].shape == (6, 6, 32, 6)
    assert int(f["optimizer_weights/Adam/iterations:0"].read()) == 5      # steps 0..4: the loop stops AFTER step == training_steps (trainer.py:417)
    capsys.readouterr()
    cli.main(common + ["--training_steps", "2", "--load_weights", out, "--save_weights", str(tmp_path / "run2.h5")])
    assert "loaded" in capsys.readouterr().out
    assert int(hdf5_lite.File(str(tmp_path / "run2.h5"))["optimizer_weights/Adam/iterations:0"].read()) == 8


def test_forward_calls_with_other_batch_sizes_keep_the_training_state():
    """the reference's own usage - model(tf.zeros([8, ...])) before training (vae/main.py:74), visualiser decodes of other
    batch sizes mid-training - must not replace the training engine: Adam moments, iteration count (staircase LR, bias correction)
    and the weights stay with it; the other batch size is served by a forward engine fed with the current weights."""
    kind, H, B = "lgvae", 32, 4
    params, batch = make_case(kind, H, B, 4, seed_base=43)
    x, eg, el = to_dev(batch["inputs"]), to_dev(batch["eps_g"]), to_dev(batch["eps_l"])
    m = _model(kind, H)
    m.set_weights_by_name(params)
    m.build(B)
    train_engine = m.engine
    for _ in range(3):
        m.engine.train_step(x, eg, el, None)
    torch.cuda.synchronize()
    v_before, p_before = m.engine.adam_v.clone(), m.engine.params.clone()
    out8 = m(torch.zeros(8, H, H, 6, device="cuda"))                 # batch 8 != training batch 4
    zx, zxh = m.encode(torch.zeros(2, H, H, 6, device="cuda"))
    rec = m.decode(torch.zeros(16, 128, device="cuda"), torch.zeros(16, 128, device="cuda"))
    torch.cuda.synchronize()
    assert out8[0].shape[0] == 8 and zx.shape[0] == 2 and rec[0].shape[0] == 16
    assert m.engine is train_engine and m.engine.B == B and m.engine.iterations == 3
    assert torch.equal(m.engine.adam_v, v_before) and torch.equal(m.engine.params, p_before)
    # the forward engine saw the TRAINED weights: decode(0) equals the training engine's own decode of zeros at its batch size
    mine = m.decode(torch.zeros(B, 128, device="cuda"), torch.zeros(B, 128, device="cuda"))[0][0].clone()
    assert torch.allclose(rec[0][0], mine, atol=1e-6)
    # an explicit rebuild for a new TRAINING batch carries the optimizer state over as well
    m.build(8)
    assert m.engine.iterations == 3 and m.engine.B == 8 and float(m.engine.adam_v.abs().sum()) == float(v_before.abs().sum())


def test_in_kernel_noise_is_fresh_on_every_forward_call():
    """forward-only calls draw new Philox noise each time (tf.random does); the counter is bumped on the device by every
    forward pass, not only by optimizer steps, and evaluation engines use their own stream."""
    kind, H, B = "lggmvae", 32, 4
    params, batch = make_case(kind, H, B, 4, seed_base=44)
    m = _model(kind, H)
    m.set_weights_by_name(params)
    x = to_dev(batch["inputs"])
    z1 = m.encode(x)[0].clone()
    y1 = m.get_y(x)[0].clone()
    z2 = m.encode(x)[0].clone()
    y2 = m.get_y(x)[0].clone()
    assert not torch.equal(z1, z2) and not torch.equal(y1, y2)
    ev = m.eval_engine(2 * B)
    ev.forward(torch.cat([x, x]), None, None, None)
    torch.cuda.synchronize()
    assert not torch.equal(ev.output("z_x")[:B], z2)


def test_checkpoint_loaded_before_the_engine_exists_restores_the_optimizer(tmp_path):
    """load_weights on a freshly constructed model (no engine yet - the normal state) must not drop the Adam moments and
    the iteration count silently: they are applied when the engine is built; a truncated checkpoint raises."""
    kind, H, B = "lggmvae", 32, 4
    params, batch = make_case(kind, H, B, 4, seed_base=45)
    x, eg, el, u = to_dev(batch["inputs"]), to_dev(batch["eps_g"]), to_dev(batch["eps_l"]), to_dev(batch["u"])
    a = _model(kind, H)
    a.set_weights_by_name(params)
    a.build(B)
    for _ in range(2):
        a.engine.train_step(x, eg, el, u)
    path = a.save_weights(str(tmp_path / "ckpt"), include_optimizer=True)
    a.engine.train_step(x, eg, el, u)
    torch.cuda.synchronize()
    b = _model(kind, H)
    b.load_weights(path)                       # no engine yet
    assert b.engine is None
    b.build(B)
    assert b.engine.iterations == 2
    b.engine.train_step(x, eg, el, u)
    torch.cuda.synchronize()
    assert torch.equal(a.engine.params, b.engine.params) and torch.equal(a.engine.adam_m, b.engine.adam_m)
    import json
    with np.load(path) as z:
        blob = {k: z[k] for k in z.files}
    assert json.loads(str(blob["keras_names"]))["encoder_x.y_dense.kernel"] == "lggm_vae/encoder/y_dense/kernel:0"
    del blob["adam_v/decoder_x.d1.kernel"]
    np.savez(str(tmp_path / "broken.npz"), **blob)
    with pytest.raises(KeyError):
        _model(kind, H).load_weights(str(tmp_path / "broken.npz"))


def test_staircase_learning_rate_step_at_one_million_iterations():
    """ExponentialDecay(lr, 1e6, 0.4, staircase=True) on optimizer.iterations (vae/main.py:67-68): the device-side schedule
    (adam_prepare) is exercised across the 1e6 boundary through sv_set_iterations and compared with the oracle's Keras Adam."""
    kind, H, B, lr = "lggmvae", 32, 2, float(np.float32(1e-3))
    params, batch = make_case(kind, H, B, 4, seed_base=46)
    x, eg, el, u = to_dev(batch["inputs"]), to_dev(batch["eps_g"]), to_dev(batch["eps_l"]), to_dev(batch["u"])
    from helpers import make_engine
    disp = {}
    for it0 in (999_999, 1_000_000, 2_000_000):
        e = make_engine(kind, H, B, "fp32", 10.0, lr=lr)
        e.load_params(params)
        e.iterations = it0
        e.train_step(x, eg, el, u)
        torch.cuda.synchronize()
        assert e.iterations == it0 + 1
        st = O.TrainState(params)
        st.iterations = it0
        O.train_step(st, kind, batch["inputs"], batch["eps_g"], batch["eps_l"], batch["u"], beta=10.0, lr=lr)
        mine = e.get_params()
        k = "decoder_x.d3.kernel"
        dm, dr = mine[k] - params[k], st.params[k] - params[k]
        assert rel_l2(dm, dr) < 0.02, (it0, rel_l2(dm, dr))
        disp[it0] = float(np.linalg.norm(dm))
    assert abs(disp[1_000_000] / disp[999_999] - 0.4) < 0.01            # the step at 1e6 (0-based iteration count)
    assert abs(disp[2_000_000] / disp[1_000_000] - 0.4) < 0.01


def test_train_metrics_accumulate_on_the_device_every_step():
    """The reference's Keras Mean metrics see EVERY train step (vae/trainer.py:140-144); here the step adds its scalars to running sums
    on the device, read and cleared at report time."""
    kind, H, B = "lgvae", 32, 4
    params, batch = make_case(kind, H, B, 4, seed_base=47)
    x, eg, el = to_dev(batch["inputs"]), to_dev(batch["eps_g"]), to_dev(batch["eps_l"])
    from helpers import make_engine
    e = make_engine(kind, H, B, "fp32", 5.0)
    e.load_params(params)
    e.output("scalar_sums").zero_()
    per_step = []
    for _ in range(4):
        e.train_step(x, eg, el, None)
        torch.cuda.synchronize()
        per_step.append(e.scalars())
    means, n = e.metric_means(reset=True)
    assert n == 4
    for k in means:
        assert abs(means[k] - np.mean([s[k] for s in per_step])) <= 1e-5 * max(1.0, abs(means[k])), k
    assert e.metric_means()[1] == 0
