"""pyqg_generative_b200 -- B200-native (sm_100a) engine for the online-simulation hot path of m2lines/pyqg_generative.

The package mirrors the reference's module names for the path it accelerates:

    pyqg_generative_b200.tools.stochastic_pyqg   EnsembleQGModel / stochastic_QGModel, AR1_sampler, constant_sampler
    pyqg_generative_b200.tools.cnn_tools         AndrewCNN, ChannelwiseScaler, apply_function
    pyqg_generative_b200.tools.operators         Operator1/2/5, cut_off, PV_subgrid_forcing
    pyqg_generative_b200.tools.simulate          run_simulation, set_initial_condition, generate_subgrid_forcing
    pyqg_generative_b200.models.*                CGANRegression, CVAERegression, MeanVarModel, OLSModel

All numerical work is done by hand-written CUDA kernels in libqgb200.so behind the C ABI of include/qgb200.h.
"""
__version__ = '0.1.0'
