"""CLI with the flags of vae/main.py:16-31 (names and defaults verbatim).  New flags have new names.

    python -m splitvae_b200.main --model lgvae --beta 120 --patch_size 8 --dataset celeba64 -no_label
"""
from __future__ import annotations

import argparse

from .utils import dotdict


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument('-viz', action='store_true')  # accepted; visualisation is out of scope
    parser.add_argument('--global_latent_dims', type=int, nargs='?', default=128)
    parser.add_argument('--local_latent_dims', type=int, nargs='?', default=128)
    parser.add_argument('--learning_rate', type=float, nargs='?', default=1e-4)
    parser.add_argument('--beta', type=float, nargs='?', default=40)
    parser.add_argument('--dataset', type=str, nargs='?', default='svhn')
    parser.add_argument('--training_steps', type=int, nargs='?', default=1000000)
    parser.add_argument('--batch_size', type=int, nargs='?', default=64)
    parser.add_argument('--patch_size', type=int, nargs='?', default=1)
    parser.add_argument('--augmentation', type=str, nargs='?', default='scramble')
    parser.add_argument('-no_label', action='store_true')
    parser.add_argument('--model', type=str, nargs='?', default='lgvae')
    parser.add_argument('--y_size', type=int, nargs='?', default=30)
    parser.add_argument('--tau', type=float, nargs='?', default=0.4)
    parser.add_argument('--alpha', type=float, nargs='?', default=40)
    parser.add_argument('-allow_growth', action='store_true')  # TF session option; no effect here
    # new in this build
    parser.add_argument('--precision', type=str, default='bf16x3', choices=['bf16x3', 'bf16', 'fp32'],
                        help='bf16x3 (default): forward on bf16 pairs, gradients within 1e-2 of fp32; bf16: single-bf16 operands, fastest; fp32: SIMT reference kernels')
    parser.add_argument('--report_every', type=int, default=10000)
    parser.add_argument('--seed', type=int, default=0)
    parser.add_argument('--no_graph', action='store_true')
    parser.add_argument('--test_batches', type=int, default=4, help='synthetic test batches per evaluation pass (0: no evaluation)')
    parser.add_argument('--data_root', type=str, default='', help='directory with the SVHN .mat files (train/extra/test_32x32.mat); default: synthetic data')
    parser.add_argument('--save_weights', type=str, default='',
                        help="write the trained weights + Adam state here: '<run>.h5' = the reference's Keras HDF5 file (vae/trainer.py:421), anything else = .npz")
    parser.add_argument('--load_weights', type=str, default='',
                        help='start from this checkpoint (.h5 Keras weights file or .npz); Adam moments and the step count resume when the file holds them')
    return parser


def make_config(argv=None):
    args = build_parser().parse_args(argv)
    config = dotdict(vars(args))
    config.label = not config.no_label           # vae/main.py:49
    config.use_graph = not config.no_graph
    return config


def main(argv=None):
    config = make_config(argv)
    print('Config:', config)
    from . import data, trainer
    from .augmentation import Augmentator
    from .model import GMVae, LGGMVae, LGVae
    augmentor = Augmentator(type=config.augmentation, size=config.patch_size, seed=config.seed)
    real = bool(config.data_root)
    train_dataset, test_dataset, input_shape = data.get_dataset(dataset=config.dataset, get_label=real and config.label,
                                                                batch_size=config.batch_size, augmentor=augmentor,
                                                                seed=config.seed, test_batches=config.test_batches,
                                                                data_root=config.data_root or None)
    if not real:
        config.label = False  # synthetic data carries no labels
    if config.model == 'lgvae':
        model = LGVae(global_latent_dims=config.global_latent_dims, local_latent_dims=config.local_latent_dims,
                      image_shape=input_shape, precision=config.precision)
        optimizer = trainer.Adam(learning_rate=config.learning_rate)
    elif config.model == 'lggmvae':
        lr_schedule = trainer.ExponentialDecay(config.learning_rate, decay_steps=1000000, decay_rate=0.4, staircase=True)
        optimizer = trainer.Adam(learning_rate=lr_schedule)
        model = LGGMVae(global_latent_dims=config.global_latent_dims, local_latent_dims=config.local_latent_dims,
                        image_shape=input_shape, y_size=config.y_size, tau=config.tau, precision=config.precision)
    elif config.model == 'gmvae':                  # vae/main.py:70-73
        lr_schedule = trainer.ExponentialDecay(config.learning_rate, decay_steps=1000000, decay_rate=0.4, staircase=True)
        optimizer = trainer.Adam(learning_rate=lr_schedule)
        model = GMVae(global_latent_dims=config.global_latent_dims, image_shape=input_shape, y_size=config.y_size, tau=config.tau,
                      precision=config.precision)
    else:
        raise NotImplementedError("--model %s: expected lgvae | lggmvae | gmvae" % config.model)
    if config.load_weights:
        model.load_weights(config.load_weights)
        print('loaded', config.load_weights)
    print('Training local-global autoencoder')
    history = trainer.train_local_global_autoencoder(model, optimizer, config.dataset, train_dataset, test_dataset, config=config)
    if config.save_weights:                       # vae/trainer.py:421
        print('saved', model.save_weights(config.save_weights, include_optimizer=True))
    return history


if __name__ == '__main__':
    main()
