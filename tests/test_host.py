"""Host-side logic on CPU: CLI surface (vae/main.py:16-31), optimizer objects, sharding, and the data-parallel
gradient exchange over gloo with world_size 2 (the N>1 path; NCCL carries the same calls on the GPU box)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import splitvae_oracle as O
from splitvae_b200 import main as cli
from splitvae_b200 import parallel, trainer
from splitvae_b200.utils import dotdict


def test_cli_defaults_match_reference_flags():
    c = cli.make_config([])
    # vae/main.py:16-31 defaults, verbatim
    assert (c.global_latent_dims, c.local_latent_dims, c.learning_rate, c.beta, c.dataset) == (128, 128, 1e-4, 40, "svhn")
    assert (c.training_steps, c.batch_size, c.patch_size, c.augmentation, c.model) == (1000000, 64, 1, "scramble", "lgvae")
    assert (c.y_size, c.tau, c.alpha, c.viz, c.no_label, c.allow_growth) == (30, 0.4, 40, False, False, False)
    assert c.label is True and c.nonexistent_key is None                   # dotdict: missing keys read as None (vae/utils.py:3-7)
    c = cli.make_config("--model lggmvae --beta 120 --alpha 40 --y_size 30 --patch_size 8 --dataset celeba64 -no_label".split())
    assert c.model == "lggmvae" and c.beta == 120 and c.patch_size == 8 and c.dataset == "celeba64" and c.label is False


def test_dataset_shapes_and_errors():
    from splitvae_b200 import data
    assert data.image_shape("svhn") == [-1, 32, 32, 3] and data.image_shape("celeba64") == [-1, 64, 64, 3]
    with pytest.raises(NotImplementedError):
        data.image_shape("mnist")                                            # vae/data.py:21


def test_optimizer_objects():
    opt = trainer.Adam(learning_rate=3e-4)
    assert opt.learning_rate == 3e-4 and opt.schedule is None
    sched = trainer.ExponentialDecay(1e-4, decay_steps=1000000, decay_rate=0.4, staircase=True)     # vae/main.py:67
    assert trainer.Adam(learning_rate=sched).learning_rate == 1e-4
    with pytest.raises(NotImplementedError):
        trainer.ExponentialDecay(1e-4, decay_steps=10, decay_rate=0.5, staircase=False)


def test_shard_range():
    assert [parallel.shard_range(2048, 8, r) for r in (0, 7)] == [(0, 256), (1792, 2048)]
    with pytest.raises(ValueError):
        parallel.shard_range(10, 4, 0)


def test_loss_functions_match_oracle_on_cpu_tensors():
    rng = np.random.default_rng(0)
    mu, sg = torch.tensor(rng.standard_normal((4, 128))), torch.tensor(rng.uniform(0.3, 2, (4, 128)))
    assert abs(float(trainer.kl_divergence(mu, sg)) - float(O.kl_divergence(mu, sg))) < 1e-12
    assert abs(float(trainer.kl_divergence_two_gauss(mu, sg, 0., 1.)) - float(O.kl_divergence_two_gauss(mu, sg, 0., 1.))) < 1e-12


def _dp_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = parallel.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    torch.set_num_threads(1)
    model, H, Bg, beta = "lgvae", 16, 4, 3.0
    params = O.init_params(model, H, H, seed=7)
    b = O.synthetic_batch(Bg, H, 2, seed_base=60)
    lo, hi = parallel.shard_range(Bg, world, rank)
    # per-rank step on its contiguous shard (the engine scales gradients by 1/(b*world); the oracle's batch mean is 1/b)
    sc, grads = O.forward_backward(params, model, b["inputs"][lo:hi], b["eps_g"][lo:hi], b["eps_l"][lo:hi], None, beta=beta,
                                   dtype=torch.float64)
    names = list(grads.keys())
    flat = torch.cat([torch.tensor(grads[k]).reshape(-1) for k in names]) / world
    split = flat.numel() // 3
    red = parallel.BucketReducer(flat, [(split, flat.numel() - split), (0, split)])      # two buckets, backward order
    assert red.enabled
    red.reduce(0)
    red.reduce(1)
    red.wait_all()
    scal = parallel.mean_scalars(torch.tensor([sc["total"], sc["recon_x"]], dtype=torch.float64))
    if rank == 0:
        ret["flat"] = flat.numpy().copy()
        ret["scal"] = scal.numpy().copy()
        ret["names"] = names
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_world_size_2_equals_global_batch():
    """N ranks x batch b with SUM all-reduce of pre-scaled gradients == one step at batch N*b."""
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29600 + os.getpid() % 300
    mp.spawn(_dp_worker, args=(2, port, ret), nprocs=2, join=True)
    model, H, Bg, beta = "lgvae", 16, 4, 3.0
    params = O.init_params(model, H, H, seed=7)
    b = O.synthetic_batch(Bg, H, 2, seed_base=60)
    sc, grads = O.forward_backward(params, model, b["inputs"], b["eps_g"], b["eps_l"], None, beta=beta, dtype=torch.float64)
    full = np.concatenate([grads[k].reshape(-1) for k in ret["names"]])
    assert np.allclose(ret["flat"], full, rtol=1e-9, atol=1e-12)
    assert np.allclose(ret["scal"], [sc["total"], sc["recon_x"]], rtol=1e-12)


def test_cli_matches_the_reference_parser():
    """splitvae_b200.main's parser against tests/golden/reference_cli.json: the reference's own argparse block (vae/main.py:15-31)
    executed verbatim by scripts/make_reference_cli_golden.py, and every `python main.py ...` command of its README parsed by it."""
    import json
    with open(os.path.join(os.path.dirname(__file__), "golden", "reference_cli.json")) as f:
        G = json.load(f)
    mine = {a.dest: a for a in cli.build_parser()._actions}
    assert len(G["flags"]) == 16
    for dest, spec in G["flags"].items():
        assert dest in mine, dest
        a = mine[dest]
        assert a.option_strings == spec["options"], dest                     # same spelling, including the single-dash flags
        assert a.default == spec["default"], dest
        assert getattr(a.type, "__name__", None) == spec["type"], dest
    assert len(G["commands"]) >= 7
    for c in G["commands"]:
        got = vars(cli.build_parser().parse_args(c["argv"]))
        for k, v in c["parsed"].items():
            assert got[k] == v, (c["argv"], k)
        cfg = cli.make_config(c["argv"])
        assert cfg.label == (not c["parsed"]["no_label"])                   # vae/main.py:49


def test_linear_assignment_matches_the_reference_source():
    """trainer.linear_assignment / CategoricalAccuracy against tests/golden/reference_cluster.json (the reference's own
    linear_assignment, vae/trainer.py:40-67, executed by scripts/make_reference_cluster_golden.py), including exact ties."""
    import json
    with open(os.path.join(os.path.dirname(__file__), "golden", "reference_cluster.json")) as f:
        G = json.load(f)
    for c in G["cases"]:
        lab = torch.tensor(c["labels"])
        labels = torch.nn.functional.one_hot(lab, c["num_class"]).float()
        pred = torch.tensor(c["pred"], dtype=torch.float64)
        out = trainer.linear_assignment(labels, pred)
        assert out.shape == labels.shape
        assert torch.argmax(out, dim=1).tolist() == c["assigned"]
        acc = trainer.CategoricalAccuracy()
        acc(labels, out)
        assert abs(acc.result() - c["accuracy"]) < 1e-12


def test_svhn_mat_reader_host_side(tmp_path):
    """data.SvhnBatches (vae/data.py:23-75 without the download): file layout X [32,32,3,N] / y [N,1], digit 0 stored as label 10 ->
    one-hot index 9, train = shuffled repeating full batches over train+extra, test = one pass with a partial last batch."""
    from scipy.io import savemat
    from splitvae_b200 import data
    rng = np.random.default_rng(0)
    sets = {}
    for name, n in (("train", 10), ("extra", 6), ("test", 7)):
        X = rng.integers(0, 256, (32, 32, 3, n), dtype=np.uint8)
        y = rng.integers(1, 11, (n, 1)).astype(np.uint8)
        y[0, 0] = 10
        savemat(str(tmp_path / f"{name}_32x32.mat"), {"X": X, "y": y})
        sets[name] = (X.transpose(3, 0, 1, 2), y.reshape(-1))
    tr = data.SvhnBatches(str(tmp_path), "train", 4, None, get_label=True, extra=True, seed=1)
    assert len(tr) == 16 and tr.shape == [-1, 32, 32, 3]
    allx = np.concatenate([sets["train"][0], sets["extra"][0]])
    ally = np.concatenate([sets["train"][1], sets["extra"][1]])
    seen = []
    for _ in range(4):                                   # one epoch = 4 full batches, every image exactly once
        u8, lab = tr.host_batch()
        assert u8.shape == (4, 32, 32, 3) and u8.dtype == torch.uint8 and lab.shape == (4, 10)
        for img, l in zip(u8.numpy(), lab.numpy()):
            j = next(i for i in range(16) if np.array_equal(allx[i], img))
            seen.append(j)
            assert np.argmax(l) == (ally[j] - 1) % 10 and l.sum() == 1
    assert sorted(seen) == list(range(16))
    u8, _ = tr.host_batch()                              # repeats: a new shuffled epoch
    assert u8.shape[0] == 4
    te = data.SvhnBatches(str(tmp_path), "test", 4, None, get_label=True)
    sizes = []
    while True:
        try:
            u8, lab = te.host_batch()
        except StopIteration:
            break
        sizes.append(u8.shape[0])
    assert sizes == [4, 3]                               # batch(B) keeps the partial last batch (vae/main.py:58)
    assert np.array_equal(te.x.numpy(), sets["test"][0])
    assert int(np.argmax(data.SvhnBatches(str(tmp_path), "test", 7, None, True).host_batch()[1][0])) == 9    # label 10 -> index 9
    nx = data.SvhnBatches(str(tmp_path), "train", 4, None, extra=False)
    assert len(nx) == 10                                 # svhn_no_extra
    with pytest.raises(FileNotFoundError):
        data.SvhnBatches(str(tmp_path / "missing"), "train", 4, None)


def test_keras_weight_name_map_covers_every_variable():
    """Checkpoint interop (vae/trainer.py:421): every variable of every model kind has a Keras HDF5 name, names are unique, layer counters
    follow the creation order of vae/model.py (conv2d .. conv2d_13, dense .. dense_5 for LGVae)."""
    from oracle import splitvae_oracle as O
    from splitvae_b200.model import keras_weight_names
    for kind, n in (("lgvae", 40), ("lggmvae", 54), ("gmvae", 34)):
        names = keras_weight_names(kind)
        ours = [f"{name}.{part}" for name, _, _, _ in O.layer_table(kind, 32, 32) for part in ("kernel", "bias")]
        assert sorted(names) == sorted(ours) and len(names) == n
        assert len(set(names.values())) == n
    lg = keras_weight_names("lgvae")
    assert lg["encoder_x.e1.kernel"] == "lg_vae/encoder/conv2d/kernel:0"
    assert lg["encoder_x_hat.e4_sd.bias"] == "lg_vae/encoder_1/dense_3/bias:0"
    assert lg["decoder_x.d1.kernel"] == "lg_vae/decoder/dense_4/kernel:0"
    assert lg["decoder_x_hat.d5.kernel"] == "lg_vae/decoder_1/conv2d_13/kernel:0"
    gm = keras_weight_names("lggmvae")
    assert gm["encoder_x.y_dense.kernel"] == "lggm_vae/encoder/y_dense/kernel:0"
    assert gm["encoder_x.z_sig.bias"] == "lggm_vae/encoder/dense_5/bias:0"
    assert gm["encoder_x_hat.e1.kernel"] == "lggm_vae/encoder_1/conv2d_3/kernel:0"
