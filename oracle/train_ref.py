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
* ``cvae_losses``                     -- models/cvae_regression.py:165-230 (forward + compute_loss, ELBO)
* ``Disc`` / ``gradient_penalty`` / ``cgan_iteration`` -- tools/cnn_tools.py:212-244 (DCGAN_discriminator, bn='None'),
  models/cgan_regression.py:173-195 and :256-292 (one iteration of train_CGAN, regression='None')
  Both pinned by tests/golden/training_cvae.npz / training_cgan.npz (unmodified reference runs).
Only tests/ (and the measurement scripts under scripts/, as the torch library baseline) may import this module.
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


# ---- CVAE (models/cvae_regression.py:165-230) ------------------------------------------------------------------------------
def cvae_losses(enc, dec, x, y, eps, decoder_var='adaptive'):
    """``CVAERegression.compute_loss`` with ymean = 0 and the reparameterisation draw ``eps`` given.  enc, dec: ``Net``."""
    res = enc(torch.cat([x, y], dim=1))
    mu, logvar = res[:, :2], res[:, 2:]
    std = torch.exp(0.5 * logvar)
    var = torch.square(std)
    yhat = dec(torch.cat([x, eps * std + mu], dim=1))
    KL = 0.5 * (torch.square(mu) + var - 1 - logvar)
    MSE = torch.square(yhat - y)
    var_p = MSE.mean().item() if decoder_var == 'adaptive' else (1. if decoder_var == 'fixed' else decoder_var)
    loss_recon = 1 / (2. * var_p) * MSE.sum(dim=(1, 2, 3)).mean()
    loss_KL = KL.sum(dim=(1, 2, 3)).mean()
    with torch.no_grad():
        var_latent = var.mean()
        extra = {'MSE': MSE.mean(), 'var_latent': var_latent, 'var_aggr': mu.var() + var_latent}
    return dict(loss=loss_recon + loss_KL, loss_recon=loss_recon, loss_KL=loss_KL, **extra)


# ---- CGAN (tools/cnn_tools.py:212-244, models/cgan_regression.py:173-195, 256-292) -----------------------------------------
class Disc(nn.Module):
    """DCGAN_discriminator(in_channels, ndf, nx, bn='None'); the Sequential indices (0, 2, 5, 8, 11) match the reference's
    state-dict keys (bn='None' puts nn.Identity at 3, 6, 9)."""

    def __init__(self, sd, nx=64):
        super().__init__()
        w = [torch.as_tensor(np.asarray(sd[k])) for k in ('0.weight', '2.weight', '5.weight', '8.weight', '11.weight')]
        def conv(t, stride, pad):
            return nn.Conv2d(t.shape[1], t.shape[0], t.shape[2], stride, pad, bias=False)
        self.net = nn.Sequential(conv(w[0], 2, 1), nn.LeakyReLU(0.2, inplace=True),
                                 conv(w[1], 2, 1), nn.Identity(), nn.LeakyReLU(0.2, inplace=True),
                                 conv(w[2], 2, 1), nn.Identity(), nn.LeakyReLU(0.2, inplace=True),
                                 conv(w[3], 2, 1), nn.Identity(), nn.LeakyReLU(0.2, inplace=True),
                                 conv(w[4], 1, 0))
        self.net.load_state_dict({k: torch.as_tensor(np.asarray(v)) for k, v in sd.items()})

    def forward(self, x):
        return self.net(x)


def gradient_penalty(D, xtrue, ytrue, yfake1, yfake2, eps, coin, lambda_gp=10):
    """cgan_regression.py:173-195 with the uniform draw ``eps`` (B,1,1,1) and the coin given."""
    if coin == 0:
        ytrue_cat = torch.cat((ytrue, yfake2.detach()), dim=1)
    else:
        ytrue_cat = torch.cat((yfake1.detach(), ytrue), dim=1)
    yfake_cat = torch.cat((yfake1.detach(), yfake2.detach()), dim=1)
    yinterp = (eps * ytrue_cat + (1 - eps) * yfake_cat).requires_grad_(True)
    Dout = D(torch.cat((xtrue, yinterp), dim=1))
    dDdy = torch.autograd.grad(outputs=Dout, inputs=yinterp, grad_outputs=torch.ones_like(Dout), retain_graph=True,
                               create_graph=True)[0].view(xtrue.shape[0], -1)
    return lambda_gp * torch.mean((torch.linalg.norm(dDdy, 2, dim=1) - 1) ** 2)


def cgan_iteration(G, D, optD, optG, x, y, z1, z2, eps, coin, g_step, lambda_drift=1e-3):
    """One pass of the loop body at cgan_regression.py:256-292 (regression='None').  G: ``Net`` 4 -> 2, D: ``Disc``; optD / optG
    may be None (gradients only).  Returns the four logged losses as floats (G_loss is None without a generator step)."""
    D.zero_grad()
    yfake1, yfake2 = G(torch.cat([x, z1], dim=1)), G(torch.cat([x, z2], dim=1))
    Dtrue1 = D(torch.cat([x, y, yfake2.detach()], dim=1))
    Dtrue2 = D(torch.cat([x, yfake1.detach(), y], dim=1))
    Dfake = D(torch.cat([x, yfake1.detach(), yfake2.detach()], dim=1))
    D_loss = -0.5 * (Dtrue1.mean() + Dtrue2.mean()) + Dfake.mean()
    D_drift = lambda_drift * (Dtrue1 ** 2).mean()
    D_grad = gradient_penalty(D, x, y, yfake1, yfake2, eps, coin)
    (D_loss + D_grad + D_drift).backward()
    if optD is not None:
        optD.step()
    G_loss = None
    if g_step:
        G.zero_grad()
        G_loss = -(D(torch.cat([x, yfake1, yfake2], dim=1))).mean()
        G_loss.backward()
        if optG is not None:
            optG.step()
    return dict(D_loss=D_loss.item(), D_grad=D_grad.item(), D_drift=D_drift.item(), G_loss=None if G_loss is None else G_loss.item())
