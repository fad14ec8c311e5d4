#!/usr/bin/env python
"""MLPResNet on MNIST with the sm_100a backend -- the counterpart of the reference's
examples/mlp_resnet/model.py (model :17-58, accuracy :61-69, epoch loop :72-97, driver
:100-154), written against ``soket_b200.api`` so that a Soket training script moves over by
changing its imports.

    python examples/mlp_resnet/train_mnist.py --data-dir examples/mlp_resnet/data      # idx-gz files
    python examples/mlp_resnet/train_mnist.py --synthetic 60000 --optimizer sgd --graph

What differs from the reference script, and why:
  * batches come from the device-resident ``DataLoader`` (one gather kernel per array per
    step instead of 2 x batch_size host->device copies, soket_b200/utils/data.py);
  * ``--train-inner`` keeps each block's inner ``Sequential`` as an attribute, so that the
    layers inside the ``Residual`` are visible to ``parameters()`` / ``train()``; without it
    the script behaves like the reference, where only the first and last ``Linear`` train
    (SURVEY.md quirk Q1);
  * ``--graph`` replays the whole training step as one CUDA-graph launch
    (soket_b200.graph.StaticStep; Adam is built with ``capturable=True`` for it).
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import soket_b200.api as soket                     # noqa: E402
from soket_b200 import nn                           # noqa: E402
from soket_b200.optim import SGD, Adam              # noqa: E402
from soket_b200.utils.data import MNIST, ArrayDataset, DataLoader   # noqa: E402

criterion = nn.SoftmaxCrossEntropyLoss()


class ResidualBlock(nn.Sequential):
    """relu(x + fn(x)), fn = Linear - norm - ReLU - Dropout - Linear - norm (model.py:17-37)."""

    def __init__(self, dim, hidden_dim=100, norm=nn.LayerNorm, drop_prob=0.01, train_inner=False):
        inner = nn.Sequential(nn.Linear(dim, hidden_dim), norm(hidden_dim), nn.ReLU(), nn.Dropout(p=drop_prob),
                              nn.Linear(hidden_dim, hidden_dim), norm(hidden_dim))
        super().__init__(nn.Residual(inner), nn.ReLU())
        if train_inner:
            self.fn = inner


class MLPResNet(nn.Sequential):
    """Linear - ReLU - num_blocks x ResidualBlock - Linear (model.py:40-58)."""

    def __init__(self, dim, hidden_dim=100, num_blocks=3, num_classes=10, norm=nn.LayerNorm, drop_prob=0.01,
                 train_inner=False):
        blocks = [ResidualBlock(hidden_dim, hidden_dim, norm, drop_prob, train_inner) for _ in range(num_blocks)]
        super().__init__(nn.Linear(dim, hidden_dim), nn.ReLU(), *blocks, nn.Linear(hidden_dim, num_classes))


def mlp_resnet_get_accuracy(Z, y):
    """Fraction of rows whose softmax arg-max equals the label (model.py:61-69): the same
    Tensor expression as the reference -- exp / max / sum / divide, argmax (int32), `==`
    against the uint8 labels, mean of the bool result in float32, one scalar read back."""
    shifted = soket.exp(Z - Z.max(-1, keepdims=True))
    probs = shifted / shifted.sum(-1, keepdims=True)
    return (probs.argmax(-1) == y).mean(dtype=soket.float32).item()


def mlp_resnet_epoch(model, dataloader, optim=None, step_graph=None):
    """One pass over `dataloader` (model.py:72-97); trains when `optim` is given.  Returns
    (mean loss, mean accuracy) over the batches, both read back every step as the reference
    does.  `step_graph` = (StaticStep, x_buf, y_buf, logits) replays a captured train step."""
    loss_sum, acc_sum = 0.0, 0.0
    for X, y in dataloader:
        if step_graph is not None:
            g, x_buf, y_buf, logits_of = step_graph
            x_buf._data[:] = X._data
            y_buf._data[:] = y._data
            g.launch()
            loss, logits = g.loss, logits_of()
        else:
            logits = model(X)
            loss = criterion(logits, y)
            if optim is not None:
                loss.backward()
                optim.step()
        loss_sum += loss.item()
        acc_sum += mlp_resnet_get_accuracy(logits, y)
    n = dataloader.max_iter
    return loss_sum / n, acc_sum / n


def synthetic_mnist(n, seed=0):
    """MNIST-shaped stand-in when the idx files are not on disk: 10 fixed class prototypes
    plus noise, so that the accuracy has something to learn."""
    protos = np.random.default_rng(1234).random((10, 784), dtype=np.float32)   # the same classes for every split
    rng = np.random.default_rng(seed)
    y = rng.integers(0, 10, n).astype(np.uint8)
    X = (0.6 * protos[y] + 0.4 * rng.random((n, 784), dtype=np.float32)).astype(np.float32)
    return X, y


def make_loaders(args):
    if args.synthetic:
        Xtr, ytr = synthetic_mnist(args.synthetic, 0)
        Xte, yte = synthetic_mnist(max(args.synthetic // 6, args.batch_size), 1)
        train, test = ArrayDataset(Xtr, ytr), ArrayDataset(Xte, yte)
    else:
        d = args.data_dir
        train = MNIST(f'{d}/train-images-idx3-ubyte.gz', f'{d}/train-labels-idx1-ubyte.gz')
        test = MNIST(f'{d}/t10k-images-idx3-ubyte.gz', f'{d}/t10k-labels-idx1-ubyte.gz')
    return (DataLoader(train, batch_size=args.batch_size, shuffle=True),
            DataLoader(test, batch_size=args.batch_size, shuffle=False))


def train_and_test_mlp_resnet_with_mnist(args):
    import soket_b200 as sk
    sk.init(0)
    train_loader, test_loader = make_loaders(args)
    np.random.seed(args.seed)
    model = MLPResNet(28 * 28, hidden_dim=args.hidden_dim, num_blocks=args.num_blocks, num_classes=10,
                      norm=nn.BatchNorm1d if args.norm == 'batch' else nn.LayerNorm,
                      drop_prob=args.drop_prob, train_inner=args.train_inner)
    for m in model.modules():
        if isinstance(m, nn.Linear):
            nn.kaiming_normal(m.weight)
    if args.optimizer == 'adam':
        optim = Adam(model.parameters(), lr=args.lr, weight_decay=args.weight_decay, capturable=args.graph)
    else:
        optim = SGD(model.parameters(), lr=args.lr, weight_decay=args.weight_decay)
    model.train(True)

    step_graph = None
    if args.graph:
        if len(train_loader.dataset) % args.batch_size:
            raise SystemExit('--graph needs a dataset size divisible by the batch size (static shapes)')
        from soket_b200.graph import StaticStep
        X0, y0 = next(iter(train_loader))
        x_buf, y_buf = soket.Tensor(X0), soket.Tensor(y0)
        holder = {}

        def one_step():
            holder['logits'] = model(x_buf)
            loss = criterion(holder['logits'], y_buf)
            loss.backward()
            optim.step()
            return loss
        g = StaticStep(one_step)
        g._graph.keep(holder['logits'])
        step_graph = (g, x_buf, y_buf, lambda: holder['logits'])

    for e in range(args.epochs):
        t0 = time.perf_counter()
        loss, acc = mlp_resnet_epoch(model, train_loader, optim, step_graph)
        dt = time.perf_counter() - t0
        print(f'Epoch: {e}, train loss: {loss}, train acc: {acc * 100} %  '
              f'({len(train_loader.dataset) / dt:.0f} samples/s)')
    model.train(False)
    loss, acc = mlp_resnet_epoch(model, test_loader)
    print(f'Test loss: {loss}, test acc: {acc * 100} %')
    return loss, acc


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('--data-dir', default=os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data'))
    ap.add_argument('--synthetic', type=int, default=0, metavar='N', help='use N synthetic MNIST-shaped samples')
    ap.add_argument('--batch-size', type=int, default=500)       # the reference's __main__ values
    ap.add_argument('--epochs', type=int, default=10)
    ap.add_argument('--optimizer', default='adam', choices=['adam', 'sgd'])
    ap.add_argument('--lr', type=float, default=0.01)
    ap.add_argument('--weight-decay', type=float, default=0.01)
    ap.add_argument('--hidden-dim', type=int, default=100)
    ap.add_argument('--num-blocks', type=int, default=3)
    ap.add_argument('--norm', default='layer', choices=['layer', 'batch'])
    ap.add_argument('--drop-prob', type=float, default=0.01)
    ap.add_argument('--train-inner', action='store_true')
    ap.add_argument('--graph', action='store_true')
    ap.add_argument('--seed', type=int, default=0)
    return train_and_test_mlp_resnet_with_mnist(ap.parse_args(argv))


if __name__ == '__main__':
    main()
