"""Command-line trainer: ``pyqg_generative/tools/train_model.py`` (the script ``scripts/train_parameterizations.py`` submits) on the
device-side trainers.

    python -m pyqg_generative_b200.tools.train_model --model CGANRegression --model_args "dict(nx=64, folder='model')" \
        --fit_args "dict(num_epochs=200)" --train_path '<dataset>/Operator2-64/*.nc'

Same flags as the reference (:11-17).  ``--train_path`` is the glob the reference hands to ``xr.open_mfdataset(..., concat_dim='run')``:
one ``<member>.nc`` per run with variables (time, lev, y, x), the layout ``tools/simulate.py --forcing yes`` writes.  Runs are split like
the reference (:39-49): the first ``nruns`` for training (tiled up to 250 when fewer are asked for), runs 250-274 for validation, 275-299
for the offline test; a dataset with fewer than 275 runs keeps its last tenth (at least one run) for validation instead.
Out of scope: ``model.test_offline`` (:56-57, the offline metric suite of models/parameterization.py:36-168 -- SURVEY section 2); the
fitted folder (weights, scalers, model_args.json, stats*.nc) is what the online path and the reference's own analysis consume.
"""
import argparse
import ast
import glob
import os

import numpy as np

from .dataset import read_netcdf


def load_runs(pattern, variables=('q', 'q_forcing_advection')):
    """Files matching ``pattern`` (sorted by their integer stem where they have one) -> dict of arrays (run, time, lev, y, x)."""
    def key(p):
        stem = os.path.splitext(os.path.basename(p))[0]
        return (0, int(stem)) if stem.isdigit() else (1, stem)
    files = sorted(glob.glob(pattern), key=key)
    if not files:
        raise FileNotFoundError('no files match %s' % pattern)
    runs = [read_netcdf(f) for f in files]
    return {v: np.stack([r[v] for r in runs]) for v in variables}


def split_runs(ds, nruns):
    """(train, validate) like the reference's ``isel(run=...)`` slices."""
    n = len(ds['q'])
    take = lambda sl: {k: v[sl] for k, v in ds.items()}
    train = take(slice(0, nruns))
    if nruns < 250 and n >= 250:
        nstacks = 250 // nruns
        train = {k: np.concatenate([v] * nstacks) for k, v in train.items()}
        print('Run dimension in training dataset: ', len(train['q']), '. Number of unique runs: ', nruns)
    if n >= 275:
        return train, take(slice(250, 275))
    nval = max(1, n // 10)
    return take(slice(0, min(nruns, n - nval))), take(slice(n - nval, n))


def main(argv=None):
    from ..models.cgan_regression import CGANRegression
    from ..models.cvae_regression import CVAERegression
    from ..models.mean_var_model import MeanVarModel
    from ..models.ols_model import OLSModel
    classes = dict(CGANRegression=CGANRegression, CVAERegression=CVAERegression, MeanVarModel=MeanVarModel, OLSModel=OLSModel)
    parser = argparse.ArgumentParser()
    parser.add_argument('--model', type=str, default='OLSModel')
    parser.add_argument('--model_args', type=str, default=str({}))
    parser.add_argument('--fit_args', type=str, default=str({}))
    parser.add_argument('--nruns', type=int, default=250)
    parser.add_argument('--train_path', type=str, required=True)
    parser.add_argument('--transfer_path', type=str, default='')     # (offline transfer test: out of scope, accepted for compatibility)
    args = parser.parse_args(argv)
    print(args)
    if args.model not in classes:
        raise ValueError('model %s is not on the accelerated path (%s)' % (args.model, ', '.join(sorted(classes))))

    def literal(text):                      # the reference eval()s these strings; dict(...) call syntax is accepted here as well
        text = text.strip()
        if text.startswith('dict(') and text.endswith(')'):
            call = ast.parse(text, mode='eval').body
            return {kw.arg: ast.literal_eval(kw.value) for kw in call.keywords}
        return dict(ast.literal_eval(text))
    train, validate = split_runs(load_runs(args.train_path), args.nruns)
    model = classes[args.model](**literal(args.model_args))
    model.fit(train, validate, **literal(args.fit_args))
    return model


if __name__ == '__main__':
    main()
