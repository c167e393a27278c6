"""CPU ORACLE (test infrastructure, NOT product code) -- restatement of the reference's regression training arithmetic.

Follows, with CPU torch fp32 modules and autograd (no reference import, so it travels to the GPU box):

* ``make_block`` / ``AndrewCNN``      -- /root/reference/pyqg_generative/tools/cnn_tools.py:79-98,125-176
  (Conv2d(padding='same', padding_mode='circular') -> ReLU -> BatchNorm2d, last block bare convolution)
* ``VarCNN``                          -- models/mean_var_model.py:14-17 (softplus on the output)
* ``compute_loss``                    -- tools/cnn_tools.py:177-182 (MSELoss)
* ``minibatch`` / ``evaluate_test`` / ``train`` -- tools/cnn_tools.py:607-700 (Adam, MultiStepLR [E/2, 3E/4, 7E/8] x 0.1)

Pinned against the reference itself: tests/golden/make_golden.py runs the *unmodified* reference ``compute_loss`` /
``train`` on a small network and commits losses, gradients, running statistics and the trained weights
(tests/golden/training.npz); tests/test_oracle_pins.py checks this restatement reproduces them.
Only tests/ may import this module.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def conv_blocks(sd):
    """[(conv index, bn index or None)] from the state-dict keys of ``AndrewCNN.conv``."""
    convs = sorted({int(k.split('.')[1]) for k, v in sd.items() if k.endswith('.weight') and np.ndim(v) == 4})
    return [(c, c + 2 if ('conv.%d.running_mean' % (c + 2)) in sd else None) for c in convs]


class Net(nn.Module):
    def __init__(self, sd, softplus=False):
        super().__init__()
        layers = []
        for c, b in conv_blocks(sd):
            w = torch.as_tensor(np.asarray(sd['conv.%d.weight' % c]))
            layers.append(nn.Conv2d(w.shape[1], w.shape[0], w.shape[2], padding='same', padding_mode='circular', bias=True))
            if b is not None:
                layers += [nn.ReLU(), nn.BatchNorm2d(w.shape[0])]
        self.conv = nn.Sequential(*layers)
        self.softplus = softplus
        self.load_state_dict({k: torch.as_tensor(np.asarray(v)) for k, v in sd.items()})

    def forward(self, x):
        y = self.conv(x)
        return F.softplus(y) if self.softplus else y

    def compute_loss(self, x, ytrue):
        return {'loss': nn.MSELoss()(self.forward(x), ytrue)}


def loss_and_grads(sd, x, y, softplus=False, dtype=torch.float32):
    """Training-mode loss, parameter gradients and the state dict after that forward (running statistics moved).
    ``dtype=torch.float64`` gives the rounding-free reference the fp32 implementations are judged against."""
    net = Net(sd, softplus).to(dtype)
    net.train()
    loss = net.compute_loss(torch.as_tensor(x).to(dtype), torch.as_tensor(y).to(dtype))['loss']
    loss.backward()
    grads = {k: p.grad.numpy().copy() for k, p in net.named_parameters()}
    return float(loss.item()), grads, {k: v.numpy().copy() for k, v in net.state_dict().items()}


def minibatch(*arrays, batch_size=64, shuffle=True):
    order = np.arange(len(arrays[0]))
    if shuffle:
        np.random.shuffle(order)
    for step in range(int(np.ceil(len(arrays[0]) / batch_size))):
        idx = order[step * batch_size:(step + 1) * batch_size]
        yield tuple(torch.as_tensor(a[idx]) for a in arrays)


def train(sd, X_train, Y_train, X_test, Y_test, num_epochs, batch_size, learning_rate, softplus=False):
    """Returns (final state dict, {'loss': [...], 'loss_test': [...]})."""
    net = Net(sd, softplus)
    net.train()
    opt = torch.optim.Adam(net.parameters(), lr=learning_rate)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[int(num_epochs / 2), int(num_epochs * 3 / 4),
                                                                   int(num_epochs * 7 / 8)], gamma=0.1)
    log = {'loss': [], 'loss_test': []}
    for epoch in range(num_epochs):
        tot, cnt = 0.0, 0
        for x, y in minibatch(X_train, Y_train, batch_size=batch_size):
            opt.zero_grad()
            loss = net.compute_loss(x, y)['loss']
            loss.backward()
            opt.step()
            tot += loss.item() * len(x)
            cnt += len(x)
        sched.step()
        log['loss'].append(tot / cnt)
        net.eval()
        tot, cnt = 0.0, 0
        for x, y in minibatch(X_test, Y_test, batch_size=batch_size):
            with torch.no_grad():
                tot += net.compute_loss(x, y)['loss'].item() * len(x)
            cnt += len(x)
        log['loss_test'].append(tot / cnt)
        net.train()
    return {k: v.numpy().copy() for k, v in net.state_dict().items()}, log
