"""Error of the tensor-core path against the fp32 reference outputs: shipped nx=48 networks + realistic q (golden), and
random-init networks on smooth / white inputs."""
import os, sys, tempfile, pathlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from conftest import golden, write_model_folder
from oracle import cnn_ref
import bench


def l2(a, b):
    return float(np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()))


def mx(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


from pyqg_generative_b200.models.cgan_regression import CGANRegression
from pyqg_generative_b200.models.cvae_regression import CVAERegression
from pyqg_generative_b200.models.mean_var_model import MeanVarModel
from pyqg_generative_b200.tools.cnn_tools import AndrewCNN
c = golden('closure_48.npz')
tmp = pathlib.Path(tempfile.mkdtemp())


class M:
    pass


m = M()
m.q, m.ny, m.nx = c['q'].astype('float64'), 48, 48
for prec in sys.argv[1:] or ['tc']:
    gan = CGANRegression(folder=write_model_folder(tmp, 'gan'), nx=48, precision=prec)
    vae = CVAERegression(folder=write_model_folder(tmp, 'vae'), precision=prec)
    gz = MeanVarModel(folder=write_model_folder(tmp, 'gz'), precision=prec)
    for name, mod, z in (('gan', gan, c['z32']), ('vae', vae, c['z32']), ('gz', gz, c['z64'])):
        y = mod.predict_snapshot(m, z)
        print('%-5s shipped %-3s: rel-L2 %.2e  max-norm %.2e' % (prec, name, l2(y, c[name + '_snapshot']), mx(y, c[name + '_snapshot'])))
    for seed in (0, 1, 2):
        sd = cnn_ref.random_state_dict(4, 2, seed=seed)
        net = AndrewCNN(4, 2, precision=prec)
        net.load_state_dict(sd)
        q = bench.synthetic_states(4, 64, seed) / np.array(bench.X_STD)[None, :, None, None]
        z = np.random.RandomState(seed).randn(4, 2, 64, 64)
        xs = torch.as_tensor(np.concatenate([q, z], 1).astype('float32'))
        xw = torch.randn(4, 4, 64, 64, generator=torch.Generator().manual_seed(seed))
        for tag, x in (('smooth q + noise', xs), ('white input', xw)):
            ref = cnn_ref.andrew_cnn_forward(sd, x, dtype=torch.float64).numpy()
            y = net(x.cuda()).cpu().numpy()
            print('%-5s random-init seed %d, %-16s: rel-L2 %.2e  max-norm %.2e' % (prec, seed, tag, l2(y, ref), mx(y, ref)))
