/* b200pinn.h -- C ABI of libb200pinn.so (hand-written sm_100a CUDA kernels).
 *
 * The reference (ZhendongS/Physics-Informed-Neural-Network-for-Explainable-Fault-
 * Diagnosis-in-Fuel-Cells) has no FFI: its hot path is a Python class surface in
 * 01_train_pinn_multiphysics_model.py ("01:" below).  Each entry point here names
 * the reference code it replaces.  The Python shims in the package bind these
 * through ctypes; nothing in a signature is a torch type.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host or the
 *    parameter is a by-value/by-pointer POD descriptor struct (host memory);
 *  - all tensors are fp32, dense, row-major; weights are [out, in] like nn.Linear;
 *  - `stream` is a cudaStream_t passed as void*; kernels never allocate or free;
 *  - return value: 0 on success, a positive cudaError_t, or a negative PINN_E_*;
 *  - one process per GPU; entry points are not re-entrant on the same workspace.
 */
#ifndef B200PINN_H
#define B200PINN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PINN_ABI_VERSION 5
#define PINN_N_IN 8        /* operating-condition features, 01:136-137 */
#define PINN_MAX_HIDDEN 8  /* hidden (tanh) layers supported */
#define PINN_N_LAMBDA 17   /* lambda_1..4, T1..5, H1..4, O1..4 (01:453-517) */

enum {
  PINN_E_ARG = -1,         /* null / inconsistent argument */
  PINN_E_SHAPE = -2,       /* unsupported layer widths */
  PINN_E_WORKSPACE = -3,   /* workspace too small */
  PINN_E_ALIGN = -4        /* pointer not 16-byte aligned */
};

/* DNN of 01:389-438: n_hidden x [Linear -> Tanh -> Dropout(p)], a linear mean head
 * `predict`, and the variance head Linear(H,H/2)-Tanh-Dropout-Linear(H/2,H/4)-Tanh-
 * Linear(H/4,1) followed by log(softplus(.)+1e-6).  width in {32,64,128,256}. */
typedef struct pinn_net {
  int32_t n_in;      /* must be PINN_N_IN */
  int32_t width;     /* H */
  int32_t n_hidden;  /* L, 1..PINN_MAX_HIDDEN */
  int32_t flags;     /* PINN_NET_* bits below: per-call options (the library keeps no process-global switches) */
  const float* W[PINN_MAX_HIDDEN]; /* layers.layer_i.weight [H, in_i]           */
  const float* b[PINN_MAX_HIDDEN]; /* layers.layer_i.bias   [H]                 */
  const float* Wp;  const float* bp;   /* predict        [1,H]     [1]          */
  const float* Wv0; const float* bv0;  /* var_layers.0   [H/2,H]   [H/2]        */
  const float* Wv1; const float* bv1;  /* var_layers.3   [H/4,H/2] [H/4]        */
  const float* Wv2; const float* bv2;  /* var_layers.5   [1,H/4]   [1]          */
} pinn_net_t;

enum { /* pinn_net_t.flags */
  PINN_NET_NO_TC_FWD = 1,   /* ablation / tests: run the 64-wide forward and MC sweep on the fp32 FFMA kernels   */
  PINN_NET_NO_TC_BWD = 2,   /* same for the 64-wide backward                                                     */
  PINN_NET_NO_WIDE_TC = 4,  /* same for the 128- / 256-wide nets (forward, MC sweep and backward)                */
  PINN_NET_PDL_NEVER = 8,   /* never chain a step's launches with programmatic dependent launch                  */
  PINN_NET_PDL_ALWAYS = 16, /* always chain them (default: only for batches of up to two tiles per SM)           */
  PINN_NET_NO_LOGVAR = 32,  /* DNN(logvar=False), 01:436: the log-variance output is identically 0, the variance
                               head receives no gradient and the aleatoric loss reduces to 0.5 * MSE             */
  PINN_NET_NO_FUSED_BWD = 64, /* ablation / tests: 64-wide backward as the two-kernel form (K2a + row table + K2b)
                               instead of the one-kernel form with on-chip weight gradients                      */
  PINN_NET_NO_WIDE_RESIDENT = 128, /* ablation / tests: 256-wide forward / MC sweep as one GEMM launch per layer
                               (activation planes in HBM) instead of the resident-activation kernel             */
  PINN_NET_NO_TMA_INPUT = 256, /* ablation / tests: the tensor-core forward / MC kernels load their input tiles with
                               plain global loads instead of TMA tensor-map copies                               */
  PINN_NET_NO_TC3 = 512     /* ablation / tests: 64-wide forward / MC sweep on the two-group 3xTF32 kernel (mlp_tc.cu)
                               instead of the three-group fp16-pair kernel (mlp_tc3.cu)                          */
};

/* Dropout control.  p == 0 means eval mode.  With masks == NULL the keep mask of
 * (sample s, pass t, dropout layer l, unit j) is drawn from Philox4x32-10 keyed by
 * `seed` with counter (s, t, l, j/8) -- identical for any GPU count or sharding; each
 * call yields eight 16-bit draws, a unit is dropped iff its draw < round(p * 2^16).
 * With masks != NULL the keep bits are read from masks[t][s][D], uint8 0/1,
 * D = L*H + H/2 (trunk layers in order, then the variance head) -- used to inject
 * the reference's own masks for parity, since the RNG streams differ.
 * Scale is 1/(1-p) as in torch (bernoulli_(1-p).div_(1-p)). */
typedef struct pinn_dropout {
  float p;
  int32_t reserved;
  uint64_t seed;
  int64_t sample_offset;   /* global index of the shard's first sample          */
  int64_t pass_offset;     /* global index of the first pass (MC) / step (train) */
  int64_t mask_sample_stride_n; /* N of the masks array (rows per pass)          */
  const uint8_t* masks;
} pinn_dropout_t;

/* Number of fp32 parameters of the DNN in canonical flat order
 * (W0,b0,...,W{L-1},b{L-1},Wp,bp,Wv0,bv0,Wv1,bv1,Wv2,bv2 == dnn.parameters()). */
int64_t pinn_param_count(int32_t width, int32_t n_hidden);

/* K1 -- DNN.forward (01:421-438): out_u[n], out_logvar[n]. */
int pinn_mlp_fwd(const pinn_net_t* net, const float* x, int64_t n,
                 const pinn_dropout_t* drop, float* out_u, float* out_logvar,
                 void* workspace, size_t workspace_bytes, void* stream);
size_t pinn_mlp_fwd_workspace_bytes(int32_t width, int32_t n_hidden, int64_t n);
/* Same for one particular call: `flags` = the pinn_net_t.flags that call will carry.  The function above covers every
 * path; for the 256-wide nets that means the per-layer GEMM form's operand planes (5 KB per sample), which the default
 * resident-activation kernel does not need (weight images + 96 B per sample). */
size_t pinn_mlp_fwd_workspace_bytes_flags(int32_t width, int32_t n_hidden, int64_t n, int32_t flags);

/* K2 -- backward of DNN.forward (autograd of 01:953) with the forward recomputed
 * in-kernel.  Upstream gradients are either given (grad_u, grad_logvar: [n]) or,
 * when both are NULL, produced in-kernel from the aleatoric loss 01:916-927 against
 * y[n] with mean over n_global samples.  grad_flat[P] receives the (unnormalised-
 * by-nothing-else) parameter gradients of this shard in canonical flat order;
 * loss_sums[4] (double) = {sum 0.5*exp(-s)(y-u)^2+0.5 s, sum |s|, sum (y-u)^2, n}. */
int pinn_mlp_bwd(const pinn_net_t* net, const float* x, int64_t n,
                 const pinn_dropout_t* drop, const float* grad_u,
                 const float* grad_logvar, const float* y, int64_t n_global,
                 float* grad_flat, double* loss_sums, void* workspace,
                 size_t workspace_bytes, void* stream);
size_t pinn_mlp_bwd_workspace_bytes(int32_t width, int32_t n_hidden, int64_t n);
/* Same for one particular call: `flags` = the pinn_net_t.flags that call will carry.  The default
 * path of the 64-wide net with 2 or 3 hidden layers needs a few MB whatever n is; the
 * PINN_NET_NO_FUSED_BWD form needs 2 KB per sample for its row table (what the function above,
 * which covers every path, reports). */
size_t pinn_mlp_bwd_workspace_bytes_flags(int32_t width, int32_t n_hidden, int64_t n, int32_t flags);

/* One train_dnn step (the loop body 01:948-955: train-mode forward, aleatoric loss, backward,
 * Adam.step with torch's default betas / eps, StepLR.step) as ONE call for a single-GPU trainer.
 * `net`'s tensors must be views into params_flat in the pinn_param_count layout; exp_avg /
 * exp_avg_sq: P floats; step_counter as in pinn_adam_step.  Same results as pinn_mlp_bwd followed
 * by pinn_adam_step (grad_scale 1, no clamp): on the tensor-core backward path the optimiser runs
 * inside the gradient-reduce launch.  grad_flat[P] still receives the gradients.
 * workspace >= pinn_mlp_bwd_workspace_bytes(). */
int pinn_train_dnn_step(const pinn_net_t* net, const float* x, int64_t n,
                        const pinn_dropout_t* drop, const float* y, int64_t n_global,
                        float* params_flat, float* exp_avg, float* exp_avg_sq,
                        int64_t* step_counter, double lr0, double gamma, int64_t step_size,
                        float* grad_flat, double* loss_sums, void* workspace,
                        size_t workspace_bytes, void* stream);
/* n_steps consecutive train_dnn steps enqueued by ONE call (same arguments; step i uses
 * drop->pass_offset + i, so drop->masks must be NULL when n_steps > 1; loss_sums = the last
 * step's sums).  Takes the host loop off the critical path when a step is ~50 us of GPU work. */
int pinn_train_dnn_steps(const pinn_net_t* net, const float* x, int64_t n,
                         const pinn_dropout_t* drop, const float* y, int64_t n_global,
                         float* params_flat, float* exp_avg, float* exp_avg_sq,
                         int64_t* step_counter, double lr0, double gamma, int64_t step_size,
                         int64_t n_steps, float* grad_flat, double* loss_sums, void* workspace,
                         size_t workspace_bytes, void* stream);
/* Data-parallel form (one process per GPU): the gradient bucket is summed over the ranks INSIDE the
 * gradient-reduce launch over NVLink peer memory (posted stores into every rank's symmetric buffer,
 * per-line flags, rank-ordered sum: bit-identical replicas), Adam + StepLR follow in the same launch and
 * all n_steps steps are enqueued by this one call.  Replaces the reference's single-process loop
 * 01:948-955 under torchrun; the only exchange SURVEY 8(e) allows on this path.  peer_buffers: device
 * array of `world` addresses, entry r = rank r's buffer of pinn_dp_bucket_words() 32-bit words, zeroed
 * once; step i carries tag first_tag + i (tags grow by one per step over the life of the buffer, > 0).
 * grad_flat may be NULL.  64-wide nets with 2..4 hidden layers; PINN_E_SHAPE otherwise. */
int64_t pinn_dp_bucket_words(int32_t width, int32_t n_hidden, int32_t world);
int pinn_train_dnn_steps_dp(const pinn_net_t* net, const float* x, int64_t n,
                            const pinn_dropout_t* drop, const float* y, int64_t n_global,
                            float* params_flat, float* exp_avg, float* exp_avg_sq,
                            int64_t* step_counter, double lr0, double gamma, int64_t step_size,
                            int64_t n_steps, const uint64_t* peer_buffers, int32_t rank,
                            int32_t world, uint32_t first_tag, float* grad_flat, double* loss_sums,
                            void* workspace, size_t workspace_bytes, void* stream);

/* K3 -- multi-physics residuals + reductions: net_f_V 01:724-765, net_f_T_simple
 * 01:869-914, net_f_T 01:767-867, net_f_H 01:621-722, net_f_O 01:535-619, the
 * mean(f^2) losses 01:1029-1034,1112,1222,1360 and their lambda-gradients. */
typedef struct pinn_scalers {
  float x_inv_scale[PINN_N_IN]; /* 1/scaler_X.scale_                             */
  float x_off[PINN_N_IN];       /* scaler_X.min_/scaler_X.scale_                 */
  float y_inv_scale, y_off;     /* same for scaler_Y (V = u*inv - off)           */
  float scale_y, min_y;         /* train_lambda's rebuilt affine 01:1017-1022    */
  float p_h2o;                  /* 10**x at Tc=55, 01:752-753 (fp32 on host)     */
  float reserved;
} pinn_scalers_t;

enum { /* families bitmask */
  PINN_FAM_V = 1, PINN_FAM_TS = 2, PINN_FAM_T = 4, PINN_FAM_H = 8, PINN_FAM_O = 16,
  PINN_FAM_DATA = 32 /* sum (y-u)^2, needs y */
};
enum { /* flags */
  PINN_RES_ACCURATE_MATH = 1, /* libdevice logf/expf/powf instead of MUFU approximations */
  PINN_RES_NO_MODE_A = 2,     /* skip the normalised-domain sums EA2/GA* (train_lambda dnn_para=True) */
  PINN_RES_NO_MODE_B = 4,     /* skip FV2/GB* (train_lambda dnn_para=False) */
  PINN_RES_NO_CLUSTER = 8     /* pinn_scalar_phase only (ablation / tests): always use the cooperative grid with its
                                 global-memory barrier, never the single thread-block cluster form */
};
enum { /* sums[] slots (double) */
  PINN_S_N = 0,
  PINN_S_FV2, PINN_S_EA2, PINN_S_DATA2,          /* sum f_V^2, sum (y-Vn)^2, sum (y-u)^2 */
  PINN_S_GA1, PINN_S_GA2, PINN_S_GA3,            /* d sum (y-Vn)^2 / d lambda_1..3       */
  PINN_S_GB1, PINN_S_GB2, PINN_S_GB3,            /* d sum f_V^2    / d lambda_1..3       */
  PINN_S_FT2, PINN_S_FTABS, PINN_S_GT1, PINN_S_GT3, PINN_S_GT5,
  PINN_S_FTE2,                                   /* sum f_T(Euler)^2                     */
  PINN_S_FH2, PINN_S_GH1, PINN_S_GH2, PINN_S_GH3, PINN_S_HACT, PINN_S_HTGT,
  PINN_S_FO2, PINN_S_GO1, PINN_S_GO2, PINN_S_GO3, PINN_S_OACT, PINN_S_OTGT,
  PINN_S_COUNT
};
enum { /* cols[] rows: cols is [PINN_C_COUNT][n] or NULL; a row is written iff its family is on */
  PINN_C_FV = 0, PINN_C_VACT, PINN_C_VOHM, PINN_C_VCONC, PINN_C_ENERNST, PINN_C_VEST5,
  PINN_C_I, PINN_C_VOUT5,
  PINN_C_FTS, PINN_C_TS_PRED, PINN_C_T_REAL,
  PINN_C_FT, PINN_C_T_PRED,
  PINN_C_FH, PINN_C_H_ACT, PINN_C_H_TGT, PINN_C_I_TOTAL,
  PINN_C_FO, PINN_C_O_ACT, PINN_C_O_TGT, PINN_C_O_Q, PINN_C_O2,
  PINN_C_COUNT
};
/* x[n,8] normalised; u[n] DNN prediction (normalised, may be NULL if no family needs
 * it); y[n] normalised labels (NULL unless FAM_DATA / mode-A sums wanted);
 * lambdas[17] device; halo_x[8]/halo_u[1]: row preceding x[0] for net_f_T when this
 * shard is not the start of the series (NULL => T_pred[0] = T_out[0], 01:857). */
int pinn_residuals(const float* x, const float* u, const float* y, int64_t n,
                   const pinn_scalers_t* scalers, const float* lambdas,
                   uint32_t families, uint32_t flags, const float* halo_x,
                   const float* halo_u, float* cols, double* sums,
                   void* workspace, size_t workspace_bytes, void* stream);
size_t pinn_residuals_workspace_bytes(int64_t n);

/* K4 -- get_MC_samples (01:1413-1491): one eval forward + T dropout passes with
 * running Welford statistics; per-pass activations are never materialised.
 * Finalised outputs (any may be NULL): pred_mean[n] = eval forward,
 * a_u[n] = sqrt(exp(mean_t logvar_t)), e_u[n] = sqrt(var_t u_t) (ddof 0).
 * Raw outputs for pass-sharded merging (any may be NULL): raw_mean[n], raw_m2[n],
 * raw_sum_logvar[n] over this call's T passes.
 * 64-wide nets cut sweeps of T >= 375 into runs of ~250 passes per tile (a function of T
 * alone: a sample's numbers never depend on the batch size or its sharding) so that
 * batches of a few tiles per SM still fill the machine; the runs' Welford triples sit in
 * `workspace` (pinn_mc_workspace_bytes: 96 B per sample) until a merge launch folds them
 * in order (Chan's update). */
int pinn_mc_dropout(const pinn_net_t* net, const float* x, int64_t n, int32_t T,
                    const pinn_dropout_t* drop, float* pred_mean, float* a_u,
                    float* e_u, float* raw_mean, float* raw_m2,
                    float* raw_sum_logvar, void* workspace, size_t workspace_bytes,
                    void* stream);
size_t pinn_mc_workspace_bytes(int32_t width, int32_t n_hidden, int64_t n);
size_t pinn_mc_workspace_bytes_flags(int32_t width, int32_t n_hidden, int64_t n, int32_t flags);   /* as above, per call */

/* f1 -- export row writer: the 22-column float64 `comprehensive_results` row of
 * create_comprehensive_results_array_v2 (01:1907-2010), incl. the segment-wise centred
 * moving average of both uncertainties (smooth_by_segments 01:1848-1872; pandas window
 * span [i-w/2, i+w/2-1], min_periods=1) and the segment labels (create_fault_labels
 * 01:2013-2031).  `cols` is the [PINN_C_COUNT][n] output of pinn_residuals; seg_ends[n_seg]
 * (device, int64) are the exclusive segment ends; segments 1..n_labeled get label = index.
 * rf_cols (optional, [n][6] doubles): a dense copy of columns 12..17 (prediction residual, the four physics
 * residuals, label) = everything the RF(t) entry points below read; feeding them this form instead of the
 * 22-column rows cuts their DRAM traffic from sparse 176-byte rows to 48 dense bytes per row. */
typedef struct pinn_export_scalers {
  double x_min[PINN_N_IN], x_scale[PINN_N_IN]; /* scaler_X.min_, scaler_X.scale_       */
  double y_min, y_scale;                       /* scaler_Y.min_[0], scaler_Y.scale_[0] */
  double min_y, scale_y;                       /* float64 affine rebuilt at 01:1920-1925 */
} pinn_export_scalers_t;
int pinn_export_rows(const float* x, const float* y, const float* pred_mean,
                     const float* a_u, const float* e_u, const float* cols,
                     const int64_t* seg_ends, int32_t n_seg, int32_t n_labeled,
                     int32_t window, const pinn_export_scalers_t* scalers, int64_t n,
                     double* out, double* rf_cols, void* stream);

/* f2 -- RF(t) risk function of 04_risk_function_early_warning_index.py, float64, for
 * n_series independent stacks laid out as results[n_series][n][row_cols] whose columns first_col .. first_col+5
 * are (res, pV, pT, pH, pO, label): row_cols = 22, first_col = 12 for comprehensive_results rows (04:58-62),
 * row_cols = 6, first_col = 0 for the compact form pinn_export_rows can emit (both even: 16-byte aligned):
 * pinn_rf_stats  = estimate_mu_sigma_normal (04:181-197): nan-mean / nan-std (ddof 1) of
 *                  columns 12..16 over label-0 rows -> mu_sigma[n_series][10] (mu, sigma);
 * pinn_rf_series = compute_rf_time_series (04:201-285) + find_first_alarm_index
 *                  (04:289-300): outputs [n_series][n]; first_alarm[n_series] = first index
 *                  with RF_smooth >= warn_threshold, -1 if none.  c_out / s_out optional. */
typedef struct pinn_rf_params {
  double z_safe, lambda_decay, k_logistic, c0_logistic, c_max, alpha_smooth, warn_threshold;
} pinn_rf_params_t;
size_t pinn_rf_workspace_bytes(int64_t n, int32_t n_series);
int pinn_rf_stats(const double* results, int64_t n, int32_t n_series, int32_t row_cols,
                  int32_t first_col, double* mu_sigma, void* workspace, size_t workspace_bytes,
                  void* stream);
int pinn_rf_series(const double* results, int64_t n, int32_t n_series, int32_t row_cols,
                   int32_t first_col, const double* mu_sigma, const pinn_rf_params_t* params,
                   double* rf_inst, double* rf_smooth, double* c_out, double* s_out,
                   int64_t* first_alarm, void* workspace, size_t workspace_bytes,
                   void* stream);

/* f3 -- torch.optim.Adam (defaults) + StepLR + box clamp, fused, state on device
 * (01:939-940,999-1002,1040-1047,...).  step_counter is int64[2] on the device, zeroed
 * once by the caller: [0] = steps taken so far (read, then incremented when
 * advance_counter != 0), [1] = internal ticket.  lr = lr0 * gamma^(step/step_size).
 * grad_scale multiplies grads first (1/N for summed gradients).  active (uint8[n],
 * optional) mirrors torch skipping parameters whose .grad is None.  lo/hi optional. */
int pinn_adam_step(float* params, const float* grads, float* exp_avg,
                   float* exp_avg_sq, int64_t n, int64_t* step_counter, double lr0,
                   double gamma, int64_t step_size, double grad_scale,
                   const uint8_t* active, const float* lo, const float* hi,
                   int32_t advance_counter, void* stream);
/* Same update for the physics scalars (n <= 32); gradients come as the double sums
 * of pinn_residuals: grad[i] = sums[grad_slot[i]] / sums[PINN_S_N]; slot < 0 => the
 * scalar gets no gradient (torch skips it) but is still clamped, as 01:1040-1047 do. */
int pinn_adam_step_from_sums(float* params, const double* sums,
                             const int32_t* grad_slot, float* exp_avg,
                             float* exp_avg_sq, int64_t n, int64_t* step_counter,
                             double lr0, double gamma, int64_t step_size,
                             const float* lo, const float* hi, void* stream);

/* A whole block of `n_steps` optimiser steps of one scalar phase in ONE cooperative launch -- the loop
 * bodies of train_lambda (01:1008-1055; families = PINN_FAM_V | PINN_FAM_DATA, flags select the physics
 * term), train_thermal (01:1107-1151; PINN_FAM_TS), train_hydrogen (01:1354-1391; PINN_FAM_H) and
 * train_oxygen (01:1204-1274; PINN_FAM_O): residual sums -> mean gradients -> Adam + StepLR + clamp,
 * repeated on the device with one barrier per step (cluster barrier for small batches, a grid
 * barrier in global memory otherwise).  Step for step it computes what
 * pinn_residuals (fast math) followed by pinn_adam_step_from_sums computes.
 * lambdas: all PINN_N_LAMBDA scalars (device, updated in place: only [first, first+count));
 * grad_slot / lo / hi: HOST arrays of `count` (<= 8) entries, meaning as in pinn_adam_step_from_sums
 * (slots must belong to the chosen family); exp_avg / exp_avg_sq: device, `count` floats;
 * sums: PINN_S_COUNT doubles, totals of the LAST step (evaluated before its update, as the
 * reference prints them).  Needs n > 0 and a device that supports cooperative launches;
 * workspace >= pinn_scalar_phase_workspace_bytes(). */
size_t pinn_scalar_phase_workspace_bytes(void);
/* Batches of up to 8 192 samples run as ONE thread-block cluster (partials in distributed shared memory,
 * hardware cluster barrier) unless `flags` carries PINN_RES_NO_CLUSTER. */
int pinn_scalar_phase(const float* x, const float* u, const float* y, int64_t n,
                      const pinn_scalers_t* scalers, float* lambdas, uint32_t families,
                      uint32_t flags, int32_t first, int32_t count,
                      const int32_t* grad_slot, const float* lo, const float* hi,
                      float* exp_avg, float* exp_avg_sq, int64_t* step_counter,
                      double lr0, double gamma, int64_t step_size, int64_t n_steps,
                      double* sums, void* workspace, size_t workspace_bytes, void* stream);

/* Data-parallel variant of pinn_adam_step: the gradient all-reduce is fused into the Adam launch over NVLink peer
 * memory (replaces `dist.all_reduce(grad)` + `optimizer.step()` of a DDP-style loop around 01:948-955).
 * peer_buffers: DEVICE array of `world` pointers, entry r = rank r's symmetric buffer laid out as
 * [64 x uint32 flags | slot 0: n floats | slot 1: n floats]; the caller has written this step's local gradient
 * bucket into slot `slot` of its own buffer (stream order); step_tag must increase by one per call on every rank. */
int pinn_adam_step_p2p(float* params, const uint64_t* peer_buffers, int32_t rank,
                       int32_t world, int32_t slot, uint32_t step_tag, float* exp_avg,
                       float* exp_avg_sq, int64_t n, int64_t* step_counter, double lr0,
                       double gamma, int64_t step_size, void* stream);

/* f4 -- the Gaussian-mixture pass behind fit_gmm_and_get_probabilities
 * (03_unsupervised_gmm_fault_diagnosis 03:360-426; sklearn GaussianMixture, covariance_type
 * "full", float64).  X[n][d] row-major (d <= 8), n_components <= 32, n_classes <= 16; all
 * pointers are device pointers.  weights[C], means[C][d], prec_chol[C][d][d] are sklearn's
 * weights_, means_, precisions_cholesky_.  One pass computes the responsibilities
 * resp = exp(log_prob - logsumexp) of sklearn's _estimate_log_prob_resp and, per request
 * (NULL = not wanted):
 *   resp[n][C]                    = gmm.predict_proba(X)                             03:392,415
 *   stats[C][1 + d + d(d+1)/2]    = sum resp, sum resp (x - mu_c), upper triangle of
 *                                   sum resp (x - mu_c)(x - mu_c)^T: the sufficient
 *                                   statistics of one EM iteration of gmm.fit       03:386-389
 *   comp_class_weight[C][K]       = sum_i resp[i][c] [labels[i] == k]  (needs labels;
 *                                   labels outside [0, K) are skipped)              03:394-405
 *   y_prob[n][K], y_pred[n]       = clip(resp @ comp_class_prob, 1e-12, 1) row-normalised
 *                                   and its first arg-max (needs comp_class_prob[C][K]) 03:415-423
 *   log_prob_norm_sum[1]          = sum_i logsumexp_c(...)  (n * gmm.score(X))
 * Reductions are fixed-order (deterministic).  workspace >= pinn_gmm_workspace_bytes(). */
size_t pinn_gmm_workspace_bytes(int32_t d, int32_t n_components, int32_t n_classes);
int pinn_gmm_pass(const double* X, int64_t n, int32_t d, int32_t n_components,
                  const double* weights, const double* means, const double* prec_chol,
                  const int32_t* labels, int32_t n_classes, const double* comp_class_prob,
                  double* resp, double* y_prob, int32_t* y_pred, double* stats,
                  double* comp_class_weight, double* log_prob_norm_sum, void* workspace,
                  size_t workspace_bytes, void* stream);

int pinn_abi_version(void);
const char* pinn_error_string(int code);
/* Device facts the host layer sizes grids with (SM count etc.). */
int pinn_device_sm_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B200PINN_H */
