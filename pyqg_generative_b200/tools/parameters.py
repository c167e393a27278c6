"""Experiment configurations: same names and values as pyqg_generative/tools/parameters.py:3-41."""


class ConfigurationDict(dict):
    """dict with copy-on-update helpers (reference parameters.py:3-32)."""

    def _update(self, d):
        out = ConfigurationDict(self)
        out.update(d)
        return out

    def nx(self, _nx):
        """Set the resolution together with the time step the reference pairs with it (parameters.py:18-29)."""
        table = {2048: 1800, 1024: 600, 512: 1800, 256: 3600, 128: 7200, 96: 7200}
        if _nx in table:
            dt = table[_nx]
        elif _nx <= 64:
            dt = 14400
        else:
            raise ValueError('no time step defined for nx=%d' % _nx)
        return self._update({'nx': _nx, 'dt': dt})


DAY = 86400
YEAR = 360 * DAY
EDDY_PARAMS = ConfigurationDict({'nx': 64, 'dt': 3600 * 4, 'tmax': 10 * YEAR, 'tavestart': 5 * YEAR})
JET_PARAMS = ConfigurationDict({'nx': 64, 'dt': 3600 * 4, 'tmax': 10 * YEAR, 'tavestart': 5 * YEAR,
                                'rek': 7e-08, 'delta': 0.1, 'beta': 1e-11})

SAMPLE_SLICE = slice(-40, None)
AVERAGE_SLICE = slice(360 * 5 * DAY, None)
AVERAGE_SLICE_ANDREW = slice(44, None)
ANDREW_1000_STEPS = 3600000
