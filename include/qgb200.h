/* qgb200.h -- C ABI of libqgb200.so, the B200-native (sm_100a) ensemble engine behind the
 * pyqg_generative online-simulation hot path.
 *
 * Every entry point takes plain pointers and sizes (no torch / numpy types).  A pointer argument flagged
 * ``on_device`` may be a CUDA device pointer (borrowed for the duration of the call, work is enqueued on
 * ``stream``) or a host pointer (the library stages it with cudaMemcpyAsync on ``stream`` and synchronises the
 * stream before returning when data flows back to the host).
 *
 * Each function cites the reference interface it replaces.  Paths are relative to the reference repository
 * (m2lines/pyqg_generative); "pyqg:" means upstream pyqg 0.7.2, the un-vendored dependency that owns the
 * dynamical-core arithmetic.
 *
 * Return value: 0 on success, a negative QGB_E* code otherwise; qgb_last_error() gives the message.
 * A handle is bound to one device and is not thread-safe (the reference is single-threaded per member,
 * pyfftw threads=1).  Numerical blow-up of a member is DATA (qgb_diag flags), never an error code.
 */
#ifndef QGB200_H
#define QGB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qgb_handle qgb_handle;

enum { QGB_OK = 0, QGB_EINVAL = -1, QGB_ECUDA = -2, QGB_ESTATE = -3, QGB_EUNSUPPORTED = -4 };

/* Model configuration = the pyqg.QGModel keyword arguments the reference passes
 * (pyqg_generative/tools/parameters.py:36-37, tools/simulate.py:118-126, pyqg: Model.__init__/QGModel.__init__)
 * plus the ensemble geometry. */
typedef struct qgb_config {
  int32_t nx;            /* grid points per side (ny == nx); 32/48/64/96 fused path, 128/256 multi-pass path */
  int32_t members;       /* ensemble members resident on this device */
  int32_t member_offset; /* global id of local member 0: Philox streams are keyed by the GLOBAL member id so a
                            run is invariant to how members are sharded over GPUs */
  int32_t device;        /* CUDA device ordinal */
  double L;              /* domain size [m]                 (pyqg default 1e6) */
  double dt;             /* time step [s]                   (tools/parameters.py:18-29 table) */
  double rek;            /* bottom drag [1/s]               (5.787e-7 eddy, 7e-8 jet) */
  double filterfac;      /* exponential filter factor       (23.6; 1e20 = sharp cut-off, tools/simulate.py:231) */
  double beta;           /* planetary vorticity gradient    (1.5e-11 eddy, 1e-11 jet) */
  double rd;             /* deformation radius [m]          (15000) */
  double delta;          /* layer thickness ratio H1/H2     (0.25 eddy, 0.1 jet) */
  double H1;             /* upper layer thickness [m]       (500) */
  double U1, U2;         /* background zonal flow [m/s]     (0.025, 0) */
} qgb_config;

/* Fill *cfg with the pyqg 0.7.2 defaults (eddy configuration at nx=64, dt=7200). */
void qgb_default_config(qgb_config* cfg);

/* pyqg: QGModel.__init__ (called at tools/stochastic_pyqg.py:78-79, tools/simulate.py:83,121,125).
 * Builds grids, inversion matrix, filter, FFT plans, state and history for cfg->members members. */
int qgb_create(const qgb_config* cfg, qgb_handle** out);
void qgb_destroy(qgb_handle* h);
const char* qgb_last_error(const qgb_handle* h); /* h may be NULL: last error of a failed qgb_create */

/* field ids for qgb_get */
enum {
  QGB_F_Q = 0,       /* double  (B,2,N,N)          potential vorticity                      (m.q)  */
  QGB_F_QH = 1,      /* complex (B,2,N,N/2+1)      its rfft2                                (m.qh) */
  QGB_F_PH = 2,      /* complex (B,2,N,N/2+1)      streamfunction, valid after qgb_invert   (m.ph) */
  QGB_F_U = 3,       /* double  (B,2,N,N)          anomaly zonal velocity, after qgb_invert (m.u)  */
  QGB_F_V = 4,       /* double  (B,2,N,N)                                                   (m.v)  */
  QGB_F_DQHDT = 5,   /* complex (B,2,N,N/2+1)      tendency used by the latest step         (m.dqhdt) */
  QGB_F_FORCING = 6, /* double  (B,2,N,N)          demeaned closure output                  (m.PV_forcing,
                                                   models/parameterization.py:23-34) */
  QGB_F_NOISE = 7,   /* float   (B,2,N,N) [gan/vae] or double (B,2,N,N) [gz]  latent noise   (m.noise_sampler.noise) */
  QGB_F_P = 8        /* double  (B,2,N,N)          streamfunction in physical space (m.p, pyqg _calc_derived_fields) */
};

/* pyqg kernel ``q`` setter (relied on at tools/operators.py:232-233, tools/simulate.py:131-132): copies q and
 * refreshes qh = rfft2(q).  Does not touch the time-stepping history.  q: double (B,2,N,N). */
int qgb_set_q(qgb_handle* h, const double* q, int on_device, void* stream);
/* pyqg: Model._initialize_time: t=0, tc=0, Adams-Bashforth restart (Euler, AB2, AB3). */
int qgb_reset_time(qgb_handle* h);
int qgb_get(qgb_handle* h, int field, void* out, int on_device, void* stream);
/* float32 copy of a real field (QGB_F_Q, _U, _V, _P): what the reference stores in its datasets (drop_vars, tools/simulate.py:
 * 16-36 converts every float64 variable to float32) converted on the device, so half the bytes cross PCIe.  With a host
 * destination and async != 0 the call returns without synchronising (pinned memory; the caller synchronises ``stream``). */
int qgb_get_f32(qgb_handle* h, int field, float* out, int on_device, int async, void* stream);
/* pyqg: PseudoSpectralKernel._invert (explicit calls tools/simulate.py:132,168; tools/operators.py:233). */
int qgb_invert(qgb_handle* h, void* stream);
/* pyqg: Model._step_forward x nsteps = _invert, _do_advection, _do_friction, _do_q_subgrid_parameterization
 * (closure evaluated on device if one is loaded), _forward_timestep.  This is the body of
 * ``for t in m.run_with_snapshots(...)`` at tools/simulate.py:137. */
int qgb_step(qgb_handle* h, int nsteps, void* stream);
int qgb_get_time(qgb_handle* h, double* t, int64_t* tc);
/* In the steady state (Adams-Bashforth level 3 reached, sampler drawing every step, nothing injected, no kernel timing, no
 * diagnostics sample due) qgb_step replays a captured CUDA graph of the step instead of launching its 3-12 kernels one by one:
 * the launch-bound regime of small ensembles.  Number of steps replayed so far by this handle;
 * the environment variable QGB_NO_GRAPH disables the replay. */
int64_t qgb_graph_replays(const qgb_handle* h);

/* ---- closure (CNN subgrid parameterization) ------------------------------------------------------------ */
enum { QGB_CLOSURE_NONE = 0, QGB_CLOSURE_GAN = 1, QGB_CLOSURE_VAE = 2, QGB_CLOSURE_GZ = 3, QGB_CLOSURE_OLS = 4,
       QGB_CLOSURE_RAW = 5 /* bare network for qgb_cnn_forward only: any cin/cout, not coupled to qgb_step */ };
enum { QGB_SAMPLER_AR1 = 0, QGB_SAMPLER_CONSTANT = 1, QGB_SAMPLER_DETERMINISTIC = 2 };
enum { QGB_PREC_FP32 = 0,  /* fp32 FFMA direct convolution (bit-for-bit deterministic, parity reference) */
       QGB_PREC_TC = 1,    /* tcgen05 implicit GEMM, fp16 split precision (<=1e-3 rel. of the fp32 reference) */
       QGB_PREC_TC_FAST = 2, /* same with a single-pass layer 2: ~18 % faster, error up to ~2e-3 with the shipped VAE/GZ nets */
       QGB_PREC_AUTO = 3   /* chosen per loaded network at its first evaluation: tc_fast if its measured relative L2 error
                              against the fp32 path is <= 7e-4 on the actual closure input, else tc if <= 1e-3, else fp32 */ };

/* One AndrewCNN (tools/cnn_tools.py:125-182) in eval mode with BatchNorm folded to a per-channel affine
 * (scale = gamma/sqrt(running_var+eps), shift = beta - running_mean*scale), laid out as torch stores it. */
typedef struct qgb_cnn_layer {
  int32_t cin, cout, ksize;
  int32_t relu_bn;       /* 1: conv -> ReLU -> affine (make_block, cnn_tools.py:79-98); 0: bare conv (last layer) */
  const float* weight;   /* (cout, cin, ksize, ksize) */
  const float* bias;     /* (cout) */
  const float* bn_scale; /* (cout) or NULL */
  const float* bn_shift; /* (cout) or NULL */
} qgb_cnn_layer;

/* Load network ``net`` (0: generator / decoder / mean net / OLS net; 1: GZ variance net) for closure ``kind``.
 * Replaces load_GAN / load_model / load_mean / load_var (models/cgan_regression.py:109-131,
 * models/cvae_regression.py:91-102, models/mean_var_model.py:82-100, models/ols_model.py:59-66).
 * Host pointers; the library packs and uploads. */
int qgb_cnn_load(qgb_handle* h, int kind, int net, int nlayers, const qgb_cnn_layer* layers);
/* ChannelwiseScaler stds (tools/cnn_tools.py:524-528 normalize/denormalize), model_weight
 * (WeightedParameterization, tools/simulate.py:242) */
int qgb_closure_config(qgb_handle* h, const float x_std[2], const float y_std[2], double weight, int precision);
/* The precision in effect (QGB_PREC_AUTO until the first evaluation has calibrated it) and, after an AUTO calibration, the
 * measured errors against the fp32 path: err = {tc rel-L2, tc max-norm, tc_fast rel-L2, tc_fast max-norm} (-1: not measured).
 * The tolerance they are held to is north_star's "parameterization output <= 1e-3 relative" (apply_function output,
 * tools/cnn_tools.py:702-735, in the relative L2 norm). */
int qgb_closure_precision(qgb_handle* h, int* precision, double err[4]);
/* stochastic_QGModel(sampling_type, nsteps) (tools/stochastic_pyqg.py:78-88); n_mean = M of predict_mean_snapshot
 * for the 'deterministic' sampler (models/cgan_regression.py:164-171, default 100). */
int qgb_set_sampler(qgb_handle* h, int kind, int nsteps, int n_mean);
int qgb_seed(qgb_handle* h, uint64_t seed);
/* Parity injection: white noise xi used by the NEXT sampler update instead of Philox (what generate_latent_noise
 * would have drawn, models/cgan_regression.py:154-155, models/mean_var_model.py:102-103).
 * dtype 0: float (B,2,N,N), 1: double (B,2,N,N).  Passing NULL clears the injection. */
int qgb_set_latent(qgb_handle* h, const void* xi, int dtype, int on_device, void* stream);
/* Parameterization.__call__(m) on the current state (models/parameterization.py:23-34): sampler update,
 * predict_snapshot, result kept as forcing (read it with QGB_F_FORCING). */
int qgb_closure_eval(qgb_handle* h, void* stream);
/* Generic pyqg QParameterization support (pyqg: Model._do_q_subgrid_parameterization; e.g. the inline ``Laplace``
 * closure at tools/simulate.py:207-225 or Weighted/Composite wrappers): install an externally computed forcing
 * dq: double (B,2,N,N) that the NEXT qgb_step adds as rfft2(dq) (used as given, not demeaned).  It is consumed by
 * one step.  Passing NULL clears it. */
int qgb_set_forcing(qgb_handle* h, const double* dq, int on_device, void* stream);
/* generate(x, z) / net.forward (models/cgan_regression.py:133-137, tools/cnn_tools.py:164-176): raw network
 * forward on caller data.  x: float (batch, cin, ny, nx) -> y: float (batch, cout, ny, nx); softplus!=0 applies
 * VarCNN's softplus (models/mean_var_model.py:14-17). */
int qgb_cnn_forward(qgb_handle* h, int net, const float* x, float* y, int batch, int ny, int nx, int softplus,
                    int precision, int on_device, void* stream);

/* ---- host-buffer convenience used by the end-to-end path ------------------------------------------------
 * set_q(host) -> nsteps -> get q(host) in one call (the reference exchanges q / dq with the host every step,
 * tools/cnn_tools.py:720-723). */
int qgb_step_host(qgb_handle* h, const double* q_in_host, double* q_out_host, int nsteps, void* stream);
/* Pipelined variant: enqueues H2D(q_in) -> rfft2 -> nsteps -> D2H(q_out) on ``stream`` and returns WITHOUT synchronising.
 * q_in / q_out must be pinned host memory and stay valid until the caller synchronises the stream.  Several handles
 * (member groups) driven on different streams overlap their PCIe transfers with each other's kernels. */
int qgb_step_host_async(qgb_handle* h, const double* q_in_host, double* q_out_host, int nsteps, void* stream);

/* ---- diagnostics ------------------------------------------------------------------------------------------
 * pyqg: QGModel._calc_ke, _calc_cfl (logged every twrite steps; assert cfl<1), per member.
 * ke, cfl: double (B); flags: int32 (B), bit0 = non-finite state, bit1 = cfl >= 1.  Any pointer may be NULL. */
int qgb_diag(qgb_handle* h, double* ke, double* cfl, int32_t* flags, int on_device, void* stream);
/* pyqg diagnostics KEspec = wv2*|ph|^2/M^2 and Ensspec = |qh|^2/M^2 summed over local members:
 * double (2, N, N/2+1) each (the accumulators that are all-reduced over NCCL by the Python layer). */
int qgb_diag_spectra(qgb_handle* h, double* kespec_sum, double* ensspec_sum, int on_device, void* stream);
/* Spectral energy budget of the current state, summed over the local members (pyqg 0.7.2
 * QGModel._initialize_model_diagnostics / Model._initialize_diagnostics lambdas, consumed by
 * tools/comparison_tools.py:91,164-189,222-263).  out: double (QGB_BUDGET_TERMS, N, N/2+1), every term / M^2:
 * KEflux, APEflux, APEgenspec, KEfrictionspec, entspec, paramspec_KEflux, paramspec_APEflux (their sum is pyqg's
 * paramspec; zero without a closure or before its first evaluation), then the enstrophy budget and the filter dissipation
 * that tools/comparison_tools.py:222-225 iterates over: ENSflux, ENSgenspec, ENSfrictionspec, Dissspec, ENSDissspec,
 * ENSparamspec.  Dissspec / ENSDissspec describe the update the next _forward_timestep performs (its Adams-Bashforth level,
 * the tendency history as it stands), which is where pyqg's _step_forward evaluates them. */
#define QGB_BUDGET_TERMS 13
int qgb_diag_budget(qgb_handle* h, double* out, int on_device, void* stream);
/* pyqg's time-averaged diagnostics (Model tavestart / taveint, _increment_diagnostics): once configured, qgb_step samples
 * KEspec (2), Ensspec (2) and the QGB_BUDGET_TERMS budget terms before every step with t >= dt, t >= tavestart and
 * tc % ceil(taveint/dt) == 0 and adds them (summed over the local members) to device accumulators.
 * qgb_diag_averages returns the accumulators, double (4 + QGB_BUDGET_TERMS, N, N/2+1), and the number of samples;
 * mean = sum / (nsamples * total members) after the cross-rank all-reduce.  reset != 0 clears them afterwards. */
int qgb_diag_config(qgb_handle* h, double tavestart, double taveint);
int qgb_diag_averages(qgb_handle* h, double* out, int64_t* nsamples, int reset, int on_device, void* stream);

/* ---- coarse-graining operators (stateless; tools/operators.py) -------------------------------------------
 * op: 1 = Operator1 (cut_off + model filter, :204-205), 2 = Operator2 (cut_off + gaussian, :207-208),
 *     4 = Operator4 (model filter of Operator2, :213-214), 5 = Operator5 (cut_off only, :216-217, :117-132).
 * in: double (batch, n, n) -> out: double (batch, nc, nc). */
int qgb_operator(int device, int op, int n, int nc, int batch, const double* in, double* out, int on_device,
                 void* stream);
/* PV_subgrid_forcing(q, nc, operator, pyqg_params, dealias) (tools/operators.py:283-287) for a batch of hi-res
 * snapshots q: double (batch,2,n,n).  dealias: 0 = 'none', 1 = '2/3-rule' (:253-257: q, u, v and the divergence low-passed with
 * pyqg's filter at filterfac = 1e20), 2 = '3/2-rule' (:258-266 through fft_interpolate to the 3n/2 grid and back; what
 * generate_subgrid_forcing uses, tools/simulate.py:90-92).  Outputs (batch,2,nc,nc) double, any may be NULL: forcing S, and the coarse model's q, u, v, psi (apply_operator_to_model, :219-236). */
int qgb_subgrid_forcing(const qgb_config* cfg, int op, int nc, int dealias, int batch, const double* q, double* forcing,
                        double* qf, double* uf, double* vf, double* pf, int on_device, void* stream);
/* fft_interpolate(x, n, N, truncate_2h=True) (tools/operators.py:134-190): spectral interpolation (N > n) or truncation
 * (N < n) of real fields.  in: double (batch, n, n) -> out: double (batch, N, N). */
int qgb_fft_interpolate(int device, int n, int N, int batch, const double* in, double* out, int on_device, void* stream);

/* Device-side timing of ONE network layer inside the running step loop (bench.py's roofline line): between
 * qgb_profile_begin and qgb_profile_end every launch of layer ``layer`` of network ``net`` is bracketed by CUDA events on
 * the launching stream.  qgb_profile_end synchronises, returns the summed duration [ms], the number of launches and
 * the images processed by them, and stops profiling. */
int qgb_profile_begin(qgb_handle* h, int net, int layer);
int qgb_profile_end(qgb_handle* h, double* total_ms, int64_t* launches, int64_t* images);
/* Same for EVERY kernel of the step (bench.py's per-kernel table; the event pairs between the back-to-back conv launches
 * cost a little overlap, so this is run outside the headline timed region).  Slots: 0-7 conv layers of network 0, 8-15 of
 * network 1, 16 spectral step kernel, 17 latent-noise kernel, 18 closure epilogue, 19 time-averaged diagnostics sampling.
 * ms / launches / units (images or members processed) are arrays of QGB_PROF_SLOTS entries, any may be NULL. */
#define QGB_PROF_SLOTS 20
int qgb_profile_all_begin(qgb_handle* h);
int qgb_profile_all_end(qgb_handle* h, double* ms, int64_t* launches, int64_t* units);

/* number of kernels launched by this library in this process (bench.py's gpu_launches) */
int64_t qgb_launch_count(void);
const char* qgb_version(void);

/* ---- training: the AndrewCNN regression trainer on the device ---------------------------------------
 * Reference: tools/cnn_tools.py:645-700 ``train(net, X_train, Y_train, X_test, Y_test, num_epochs, batch_size, learning_rate)``
 * -- Adam (torch defaults), ``net.compute_loss`` = MSELoss (:177-182), BatchNorm2d in training mode -- as used by
 * models/mean_var_model.py:41-66 (GZ two-stage fit: AndrewCNN for the mean, VarCNN = softplus(AndrewCNN) on the squared
 * residuals) and models/ols_model.py ``fit``.  A trainer owns parameters, Adam moments, BatchNorm running statistics and
 * activations for minibatches of up to max_batch images of ny x nx; fp32 arithmetic like the reference.
 * Network: nlayers convolutions (circular 'same', bias), channels[0..nlayers], kernel sizes ksizes[0..nlayers-1] in {1,3,5};
 * ReLU + BatchNorm2d after every layer but the last (tools/cnn_tools.py:79-98,125-160); softplus != 0 applies softplus to
 * the output (VarCNN, mean_var_model.py:14-17).
 * Flat parameter layout (qgb_train_num_params floats), per layer in order: conv weight (cout,cin,k,k), conv bias (cout),
 * [BatchNorm weight (cout), BatchNorm bias (cout)] -- the order of ``net.parameters()``.  Flat buffer layout
 * (qgb_train_num_buffers floats), per BatchNorm layer: running_mean (cout), running_var (cout). */
typedef struct qgb_trainer qgb_trainer;
int qgb_train_create(int device, int nlayers, const int32_t* channels, const int32_t* ksizes, int ny, int nx, int max_batch,
                     int softplus, qgb_trainer** out);
void qgb_train_destroy(qgb_trainer* t);
const char* qgb_train_last_error(const qgb_trainer* t); /* t may be NULL: last error of a failed qgb_train_create */
int64_t qgb_train_num_params(const qgb_trainer* t);
int64_t qgb_train_num_buffers(const qgb_trainer* t);
int64_t qgb_train_launch_count(const qgb_trainer* t);   /* kernels launched by this trainer so far */
/* host pointers (either may be NULL = unchanged); reset_optimizer != 0 clears the Adam moments and step count
 * (``optim.Adam(net.parameters(), lr)`` is created anew by every ``train`` call, cnn_tools.py:671) */
int qgb_train_set_params(qgb_trainer* t, const float* params, const float* buffers, int reset_optimizer);
int qgb_train_get_params(qgb_trainer* t, float* params, float* buffers);
/* One iteration of the loop at cnn_tools.py:685-690: optimizer.zero_grad(); loss = MSE(net(x), y); loss.backward();
 * optimizer.step() with learning rate lr.  x: float (batch, channels[0], ny, nx), y: float (batch, channels[nlayers], ny, nx),
 * host pointers (copied on ``stream``) or device pointers (on_device != 0).  *loss (host, may be NULL) = the minibatch loss. */
int qgb_train_step(qgb_trainer* t, const float* x, const float* y, int batch, int on_device, double lr, double* loss, void* stream);
/* Forward (training mode) + backward without the optimizer step: grads (host, flat parameter layout) and the loss -- the hook
 * the parity tests compare with torch autograd.  update_running == 0 leaves the BatchNorm running statistics untouched. */
int qgb_train_grads(qgb_trainer* t, const float* x, const float* y, int batch, int on_device, float* grads, double* loss,
                    int update_running, void* stream);
/* ``evaluate_test`` (cnn_tools.py:624-643): eval-mode (running statistics) MSE of one minibatch. */
int qgb_train_eval_loss(qgb_trainer* t, const float* x, const float* y, int batch, int on_device, double* loss, void* stream);
/* gradients left by the last backward pass (host, flat parameter layout) */
int qgb_train_get_grads(qgb_trainer* t, float* grads);
/* Adam betas of this trainer (torch defaults (0.9, 0.999); train_CGAN uses (0.5, 0.999), cgan_regression.py:246-247) */
int qgb_train_set_adam(qgb_trainer* t, double beta1, double beta2);

/* ---- CVAE: one iteration of train_CVAE (models/cvae_regression.py:283-289) ------------------------------------------
 * optimizer.zero_grad(); losses = net.compute_loss(x, y, ymean = 0); losses['loss'].backward(); optimizer.step() with
 * Adam over chain(encoder, decoder) (:268).  enc: AndrewCNN 4 -> 4 on cat[x, y] giving [mu, logvar] (:104-112); dec:
 * AndrewCNN 4 -> 2 on cat[x, z], z = eps exp(logvar / 2) + mu (:165-175); loss = sum (yhat - y)^2 / (2 var_p B) +
 * sum 0.5 (mu^2 + var - 1 - logvar) / B (:177-230).  x, y, eps: float (batch, 2, ny, nx), eps = the standard normal draw
 * of ``torch.randn_like(std)``; host pointers (copied on ``stream``) or device pointers (on_device != 0).
 * decoder_var < 0: 'adaptive' (var_p = the batch's mean squared error, held constant in the gradient), otherwise the value
 * ('fixed' = 1).  update != 0: Adam step of both networks with learning rate lr; 0: gradients only (qgb_train_get_grads).
 * losses (host, 6 doubles, may be NULL) = loss, loss_recon, loss_KL, MSE, var_latent, var_aggr. */
int qgb_train_cvae_step(qgb_trainer* enc, qgb_trainer* dec, const float* x, const float* y, const float* eps, int batch,
                        int on_device, double lr, double decoder_var, int update, double* losses, void* stream);

/* ---- CGAN: the discriminator and one iteration of train_CGAN (models/cgan_regression.py:227-300) -------------------
 * qgb_disc = DCGAN_discriminator(in_channels = 6, ndf, nx, bn = 'None') (tools/cnn_tools.py:212-244): four 4 x 4 / stride 2 /
 * zero-pad 1 convolutions (in_channels -> ndf -> 2 ndf -> 4 ndf -> 8 ndf) each followed by LeakyReLU(0.2), then an
 * (nx/16) x (nx/16) valid convolution to one number per sample; no biases.  Flat parameter layout: the five weights
 * (cout, cin, k, k) in order -- the order of ``D.parameters()`` (state_dict keys 0.weight, 2.weight, 5.weight, 8.weight,
 * 11.weight).  Adam with betas (0.5, 0.999) (:246).  max_batch = the minibatch size B (the object holds 4 B samples).
 * Arithmetic: every convolution is one GEMM on an im2col matrix; the GEMMs run on the tensor cores (tcgen05 kind::tf32 with the 3-term
 * split a b ~= a_hi b_hi + a_hi b_lo + a_lo b_hi, fp32 accumulation in TMEM: 1e-5 of float64 on the forward pass, csrc/tgemm.cuh) or,
 * with QGB_DISC_GEMM=ffma in the environment, in fp32 FFMA (2e-6). */
typedef struct qgb_disc qgb_disc;
int qgb_disc_create(int device, int in_channels, int ndf, int nx, int max_batch, qgb_disc** out);
void qgb_disc_destroy(qgb_disc* d);
const char* qgb_disc_last_error(const qgb_disc* d);     /* d may be NULL: last error of a failed qgb_disc_create */
int64_t qgb_disc_num_params(const qgb_disc* d);
int64_t qgb_disc_launch_count(const qgb_disc* d);
int qgb_disc_set_params(qgb_disc* d, const float* params, int reset_optimizer);        /* host pointer */
int qgb_disc_get_params(qgb_disc* d, float* params, float* grads);                      /* host pointers, either may be NULL */
/* D(x): x float (batch, 6, nx, nx), batch <= 4 max_batch -> out float (batch) */
int qgb_disc_forward(qgb_disc* d, const float* x, int batch, int on_device, float* out, void* stream);
/* One iteration of the loop at cgan_regression.py:256-292 (regression = 'None'):
 *   yfake1 = G(x, z1), yfake2 = G(x, z2) (training mode: batch statistics, two running-statistics updates);
 *   D_loss = -0.5 (mean D(x, y, yfake2) + mean D(x, yfake1, y)) + mean D(x, yfake1, yfake2);  D_drift = 1e-3 mean D(x, y, yfake2)^2;
 *   D_grad = 10 mean_b (|dD/dy (x, yinterp)|_2 - 1)^2 with yinterp = eps ytrue_cat + (1 - eps) yfake_cat, ytrue_cat = (y, yfake2)
 *   if coin == 0 else (yfake1, y) (gradient_penalty :173-195; its second-order term is evaluated exactly: D is piecewise
 *   linear, so d D_grad / d W = the weight gradient of the linearised network driven by d D_grad / d(dD/dy));
 *   (D_loss + D_grad + D_drift).backward(); optimizerD.step()  [update_d != 0, learning rate lr_d];
 *   g_mode 1 / 2: G_loss = -mean D(x, yfake1, yfake2) with the discriminator as it now is, backward through D and both
 *   generator passes; g_mode 2 also applies optimizerG.step() (lr_g) -- the reference does this every 5th iteration (:277).
 * G: qgb_trainer of the generator AndrewCNN 4 -> 2 (set its betas with qgb_train_set_adam(G, 0.5, 0.999)).
 * x, y, z1, z2: float (batch, 2, ny, nx), host or device (on_device); eps: HOST array of batch floats (torch.rand(B,1,1,1), :176);
 * coin: np.random.randint(0, 2) (:178).  losses (host, 4 doubles, may be NULL) = D_loss, D_grad, D_drift, G_loss (G_loss only
 * when g_mode != 0).  Gradients: qgb_disc_get_params(.., grads) and qgb_train_get_grads(G, ..). */
int qgb_train_cgan_step(qgb_trainer* G, qgb_disc* D, const float* x, const float* y, const float* z1, const float* z2,
                        const float* eps, int coin, int batch, int on_device, double lr_d, double lr_g, int update_d, int g_mode,
                        double* losses, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QGB200_H */
