/*
 * odpd.h — C ABI of libodpd.so: B200-native (sm_100a) recurrent-backbone forward / backward for OpenDPD.
 *
 * This is the drop-in boundary for ONE hot path of lab-emi/OpenDPD: the backbone forward/backward (+ I/Q MSE)
 * that `net_train` executes per step (reference: modules/train_funcs.py:28-48 -> models.py:150-176 ->
 * backbones/<name>.py forward).  The reference has no FFI of its own (it is pure PyTorch); the entry points
 * below are what a ctypes/cffi binding added to the reference's backbones/<name>.py would call — one call replaces
 * one `Backbone.forward` (or its autograd backward).  INTEGRATION.md shows that binding.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer owned by the caller (PyTorch), valid until the
 *     stream work completes; the library keeps no reference after the call returns and never allocates
 *     device memory.  All tensors are contiguous fp32 unless stated.
 *   - launches are enqueued on `stream` (a cudaStream_t passed as void*); no host or device synchronisation.
 *   - return 0 on success, <0 on error; odpd_last_error() returns a thread-local message. No C++ exceptions
 *     cross the ABI and the library never calls exit().
 *   - NaN/Inf propagate exactly as IEEE arithmetic dictates (the reference divides by amp=0 unguarded,
 *     backbones/dgru.py:66-67).
 *
 * Flat parameter layout (`params`, `gparams`): the reference module's named_parameters() order, each tensor
 * row-major, concatenated (H hidden, F features, K dvr units):
 *   GRU       (gru.py:17-25)        rnn.weight_ih_l0(3H,2) weight_hh_l0(3H,H) bias_ih_l0(3H) bias_hh_l0(3H) fc_out.weight(2,H) fc_out.bias(2)
 *   LSTM      (lstm.py:17-25)       rnn.weight_ih_l0(4H,2) weight_hh_l0(4H,H) bias_ih_l0(4H) bias_hh_l0(4H) fc_out.weight(2,H) fc_out.bias(2)
 *   DGRU      (dgru.py:22-33)       rnn.* as GRU with F=6, fc_out.weight(2,H+6) fc_out.bias(2) fc_hid.weight(H,H) fc_hid.bias(H)
 *   QGRU/QGRU_AMP1 (qgru.py:22-31)  rnn.* as GRU with F=4, fc_out.weight(2,H) fc_out.bias(2)
 *   DELTAGRU  (deltagru.py:23-30)   rnn.weight_ih_l0(3H,6) weight_hh_l0(3H,H) bias_ih_l0(3H) bias_hh_l0(3H) fc_out.weight(2,H) fc_out.bias(2)
 *   TRES      (deltagru_tcnskip.py:24-49) rnn.x2h.weight(3H,6) rnn.h2h.weight(3H,H) fc_out.weight(2,H) tcn.0.weight(3,2,3) tcn.2.weight(2,3,1)
 *   PGJANET   (pgjanet.py:13-22)    W_a.weight(H,H+1) W_a.bias(H) W_p1.* W_p2.* W_f.weight(H,2H) W_f.bias(H) W_g.* W_o.weight(2,H) W_o.bias(2)
 *   DVRJANET  (dvrjanet.py:14-30)   cs(K) W_ph.weight(H,H) W_ptheta.weight(H,1) W_ah.weight(H,H) W_ax.weight(H,1) W_f.weight(H,H) W_f.bias(H)
 *                                   W_ccos.weight(H,2H) W_ccos.bias(H) W_csin.* W_o1.weight(1,H) W_o1.bias(1) W_o2.weight(1,H) W_o2.bias(1)
 *   GMP       (gmp.py:11)           Weight(1,495)
 *   VDLSTM    (vdlstm.py:28-41)     rnn.weight_ih_l0(4H,4) weight_hh_l0(4H,H) bias_ih_l0(4H) bias_hh_l0(4H) fc_lambda_1.weight(4,H) .bias(4)
 *                                   fc_lambda_2.weight(4,H) .bias(4) fc_out.weight(2,8) fc_out.bias(2)
 *   RVTDCNN   (rvtdcnn.py:20-33)    Conv2d.weight(3,1,3,3) Conv2d.bias(3) fc_hid.weight(H,36) fc_hid.bias(H) fc_out.weight(2,H) fc_out.bias(2)
 *   BOJANET   (bojanet.py:14-26)    fir_I.weight(6,16) fir_Q.weight(6,16) W_fi.weight(H,12) W_fi.bias(H) W_fh.weight(H,H) W_gi.weight(H,12) W_gi.bias(H)
 *                                   W_gh.weight(H,H) W_out_I.weight(1,H) W_out_I.bias(1) W_out_Q.weight(1,H) W_out_Q.bias(1)
 *   TCNN      (tcnn.py:15-31)       network.0.weight(H,6,1) network.0.bias(H) network.{2,4,6,8}.weight(H,1,5) network.10.weight(2,H,1)
 *   NEURALTX  (neuraltx.py:19-38)   conv_I.weight(1,1,5) conv_Q.weight(1,1,5) network.0.weight(H,4,1) network.0.bias(H) network.{2,4,6,8}.weight(H,1,5)
 *                                   network.10.weight(2,H,1) IQ_match.weight(2,2)
 *   APNRRU    (apnrru.py:44-51,13-19) fir_I.weight(3,16) fir_Q.weight(3,16) rru.C(1) rru.Z(1,S) rru.W_u.weight(16,S+8) rru.W_u.bias(16) rru.W_h.weight(S,16)
 *                                   rru.W_h.bias(S) output_layer_I.weight(1,H) output_layer_Q.weight(1,H),  S = 2H+3
 *   MCLDNN    (mcldnn.py:21-27)     conv2d_1.weight(H,1,3,3) .bias(H) conv1d.weight(5H,1,3) .bias(5H) conv2d_2.weight(1,10,3,3) .bias(1) lstm.weight_ih_l0(32,5H)
 *                                   weight_hh_l0(32,8) bias_ih_l0(32) bias_hh_l0(32) fc_out.weight(16,8) .bias(16) fc_out_2.weight(2,16) .bias(2)
 *   DELTAJANET (deltajanet.py:97-105,27-29) rnn.weight_ih_l0(2H,6) weight_hh_l0(2H,H) bias_ih_l0(2H) bias_hh_l0(2H) fc_out.weight(2,H) fc_out.bias(2)
 *   TRES_QAT  (quant_envs.py:286-305 applied to deltagru_tcnskip.py) rnn.x2h.weight(3H,6) + weight/act/out scales | rnn.h2h.weight(3H,H) + 3 scales |
 *                                   rnn.add / mul / sigmoid / tanh .quantizer.scale | fc_out.weight(2,H) + 3 scales | tcn.0.weight(3,2,3) tcn.2.weight(2,3,1)
 *   QGRU_QAT  (quant_envs.py:215-305 applied to qgru.py) rnn.rnn_cell_list.0.x2h.weight(3H,4) .bias(3H) .weight_quantizer.scale .act_quantizer.scale
 *                                   .out_quantizer.scale | h2h.weight(3H,H) .bias(3H) + 3 scales | sigmoid/tanh/add/mul .quantizer.scale |
 *                                   fc_out.weight(2,H) .bias(2) + 3 scales.   For the QAT cells OdpdDims.K packs n_bits_w | n_bits_a<<8 | eval<<16.
 *   num_layers = L > 1 (GRU, LSTM, DGRU, QGRU, QGRU_AMP1; OdpdDims.K = L): the rnn.* block repeats per layer in nn.RNNBase's order —
 *                                   weight_ih_l{k}(G*H, F for k = 0 else H) weight_hh_l{k}(G*H,H) bias_ih_l{k}(G*H) bias_hh_l{k}(G*H), k = 0..L-1 —
 *                                   followed by the head tensors as above.
 *
 * Two implementations sit behind the GRU / LSTM / DGRU / QGRU entries: the fused warp-specialised kernels (H <= 32, one layer: every
 * script of record) and a layered path (H 33..64 and/or stacked layers: time-parallel projections around a per-layer chain kernel,
 * not time-chunked).  The library picks by (H, K); the calls, buffers and results have the same meaning.
 */
#ifndef ODPD_H_
#define ODPD_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ODPD_VERSION 100

/* cell families == reference backbones (models.py:26-141 names them) */
enum {
    ODPD_CELL_GRU = 0,       /* backbones/gru.py:45-48 */
    ODPD_CELL_LSTM = 1,      /* backbones/lstm.py:45-48 */
    ODPD_CELL_DGRU = 2,      /* backbones/dgru.py:59-74 */
    ODPD_CELL_DELTAGRU = 3,  /* backbones/deltagru.py:60-77, 208-266 */
    ODPD_CELL_TRES = 4,      /* backbones/deltagru_tcnskip.py:89-103, 244-304 */
    ODPD_CELL_PGJANET = 5,   /* backbones/pgjanet.py:24-77 */
    ODPD_CELL_DVRJANET = 6,  /* backbones/dvrjanet.py:43-102 */
    ODPD_CELL_GMP = 7,       /* backbones/gmp.py:18-51 */
    ODPD_CELL_QGRU = 8,      /* backbones/qgru.py:59-71 */
    ODPD_CELL_QGRU_AMP1 = 9, /* backbones/qgru_amp1.py:59-76 */
    ODPD_CELL_QGRU_QAT = 10, /* qgru.py under --quant: quant/modules/gru.py:32-124 + quant/qmodules (fake-quant QAT) */
    ODPD_CELL_QGRU_AMP1_QAT = 11, /* qgru_amp1.py under --quant */
    ODPD_CELL_VDLSTM = 12,   /* backbones/vdlstm.py:58-82; frame_length >= 3 (the window wraps over the last 3 samples) */
    ODPD_CELL_RVTDCNN = 13,  /* backbones/rvtdcnn.py:36-62 (row f-4): H = fc_hid_size (1..64); frame_length >= 3 */
    ODPD_CELL_BOJANET = 14,  /* backbones/bojanet.py:54-106 (row f-4): hidden_size 1..18 (the reference's pr_block covers 3 x 6 units); frame_length >= 15 */
    ODPD_CELL_TCNN = 15,     /* backbones/tcnn.py:83-97 (row f-4): H = hidden_channels (1..64) */
    ODPD_CELL_NEURALTX = 16, /* backbones/neuraltx.py:107-124 (row f-4): H = hidden_channels (1..64) */
    ODPD_CELL_APNRRU = 17,   /* backbones/apnrru.py:52-135 (row f-4): hidden_size 1..14 (2H+3 state values, one per warp lane); frame_length >= 15 */
    ODPD_CELL_MCLDNN = 18,   /* backbones/mcldnn.py:83-113 (row f-4): H = conv channels (1..12); the LSTM inside is always 8 wide; frame_length >= 4 */
    ODPD_CELL_DELTAJANET = 19, /* backbones/deltajanet.py:49-60, 203-262 (row f-4): hidden_size 1..16; thx / thh are ignored like in the reference (:22-26) */
    ODPD_CELL_TRES_QAT = 20, /* deltagru_tcnskip.py under --quant (quant/quant_envs.py:286-305; bash_scripts/OpenDPDv2.sh:47-49): hidden_size 1..16,
                                K packs n_bits_w | n_bits_a<<8 | eval<<16 like the QAT GRU cells; thx / thh / stats as for ODPD_CELL_TRES */
    ODPD_CELL_COUNT = 21
};

/* flags */
#define ODPD_F_NEED_DX 1u    /* backward emits gx   (frozen-PA input gradient, models.py:169-171) */
#define ODPD_F_NEED_DW 2u    /* backward emits gparams */
#define ODPD_F_SAVE 4u       /* forward stores activations for a later odpd_backbone_bwd */
#define ODPD_F_OVERWRITE_DW 8u /* backward: gparams = sum instead of gparams += sum (saves the caller a memset) */
#define ODPD_F_ZERO_LOSS 16u /* forward: the library clears *loss (cudaMemsetAsync on `stream`) before accumulating */
#define ODPD_F_X_BF16 32u      /* `x` holds bf16 pairs instead of fp32 pairs (storage only: widened exactly, arithmetic stays fp32) */
#define ODPD_F_TARGET_BF16 64u /* same for `target` */

typedef struct OdpdDims {
    int32_t cell;   /* ODPD_CELL_* */
    int32_t B;      /* sequences in this call (>=0)                        */
    int32_t T;      /* frame length (>=0)                                  */
    int32_t H;      /* hidden size: 1..32 on the fused kernels; GRU/LSTM/DGRU/QGRU/QGRU_AMP1 also 33..64 (layered kernels); ignored by GMP */
    int32_t K;      /* DVRJANET num_dvr_units (dvrjanet.py:6); QAT cells: bit widths; GRU/LSTM/DGRU/QGRU/QGRU_AMP1: num_layers
                       (0 or 1 = one layer, up to 8: nn.GRU / nn.LSTM stacking, gru.py:17-24, arguments.py:51,60) */
    uint32_t flags; /* ODPD_F_*                                            */
    float thx, thh; /* delta thresholds (deltagru.py:216-217)              */
    int32_t tchunks; /* GRU/DGRU/QGRU/LSTM/PGJANET/DVRJANET: time chunks a sequence is cut into and run concurrently (see below).
                        0 = the library picks from B, T and the SM count; 1 = plain serial recurrence; n = exactly n (<=32) */
    int32_t twarm;   /* warm-up steps in front of every chunk (rounded up to 32); 0 = the cell's default (128; JANET cells 256) */
    /* On-device framing (replaces IQFrameDataset's materialised stride-1 frames, modules/data_collector.py:233-252, and the
     * per-step H2D copy of them, train_funcs.py:30-31): when x_starts != NULL, `x` is the raw (N,2) IQ stream and sequence b
     * reads samples x_starts[b] .. x_starts[b]+T-1 of it (device int32[B]); same for target_starts / `target`.  NULL = `x` /
     * `target` are framed (B,T,2) tensors.  out, gout, gx and the saved activations are always framed. */
    const int32_t *x_starts;
    const int32_t *target_starts;
} OdpdDims;

/*
 * Time-chunked execution (GRU-family cells).  The reference's frames are T-step serial recurrences (nn.GRU inside
 * gru.py:46 / dgru.py:70 / qgru.py:69) and its batches hold fewer sequences than a B200 has SMs.  A GRU forgets its initial
 * state geometrically, so the kernels cut each sequence into `tchunks` chunks that run concurrently, each preceded by `twarm`
 * warm-up steps started from h = 0 (backward: from dL/dh = 0, in reverse time); a verify pass then compares the state every
 * chunk was started from with the state its predecessor really ended with (max|diff| <= 2^-18 * max|state| at the boundary)
 * and re-runs, serially, every sequence with a failing boundary.  Results therefore never depend on the forgetting
 * assumption; only the speed does.  odpd_chunk_plan reports what a call with these dims will do:
 *   out[0] chunks, out[1] steps per chunk, out[2] warm-up steps, out[3] index (in 4-byte units, or -1) of an int32 counter
 *   inside `saved` (backward = 0) / `workspace` (backward = 1) that counts sequences the verify pass had to re-run, followed by
 *   a float holding the largest boundary mismatch seen so far in units of the tolerance; the caller zeroes both when it
 *   allocates the buffer.  LSTM, PGJANET and DVRJANET are chunked the same way (JANET default warm-up 256); the delta cells,
 *   GMP and the QAT cell always run serially.
 */
int odpd_chunk_plan(const OdpdDims *d, int32_t backward, int32_t out[4]);
/* The planner behind odpd_chunk_plan as a pure host function (no CUDA call; unit-testable without a GPU): the plan for B sequences
 * of T steps when the device holds `slots` CTAs of the kernel at once, tchunks / twarm as in OdpdDims, default_warm = the cell's
 * default warm-up.  out[0] chunks, out[1] steps per chunk, out[2] warm-up steps. */
int odpd_chunk_plan_model(int32_t B, int32_t T, int32_t tchunks, int32_t twarm, int32_t slots, int32_t default_warm, int32_t out[3]);

int odpd_version(void);
const char *odpd_last_error(void);

/* number of fp32 parameters of a backbone == reference count_net_params (utils/util.py) of `backbone` */
int64_t odpd_n_params(int32_t cell, int32_t H, int32_t K);
/* bytes of the caller-allocated `saved` buffer odpd_backbone_fwd needs for these dims: the activations a forward with
 * ODPD_F_SAVE stores for the matching backward, plus (GRU-family, tchunks != 1) a small chunk scratch that is also needed
 * without ODPD_F_SAVE (0 = no buffer needed, `saved` may be NULL) */
int64_t odpd_saved_bytes(const OdpdDims *d);
/* bytes of the caller-allocated scratch `workspace` odpd_backbone_bwd needs (per-(sequence,chunk) grad partials + chunk scratch) */
int64_t odpd_bwd_workspace_bytes(const OdpdDims *d);

/*
 * Forward: replaces `backbone.forward(x, h_0)` with h_0 == 0 (models.py:154-155) and, when `target` is given,
 * the `nn.MSELoss()(out, target)` that follows it (train_funcs.py:35-37, project.py:262-272).
 *   x       (B,T,2)   in
 *   target  (B,T,2)   in, or NULL
 *   params  flat      in  (layout above; must be 16-byte aligned)
 *   out     (B,T,2)   out
 *   loss    double[1] in/out, or NULL: += sum((out-target)^2) * loss_scale  (caller zeroes it; loss_scale is
 *                     1/(2*B_global*T) for MSELoss 'mean')
 *   saved   odpd_saved_bytes(d) out when ODPD_F_SAVE, else may be NULL
 *   stats   int64[4]  in/out, or NULL: += {dx_zeros, dx_numel, dh_zeros, dh_numel} (delta cells only;
 *                     deltagru.py:241-247 keeps these in fp32 tensors, we count exactly in int64)
 */
int odpd_backbone_fwd(const OdpdDims *d, const float *x, const float *target, const float *params, float *out,
                      double *loss, double loss_scale, void *saved, int64_t *stats, void *stream);

/*
 * Backward: replaces autograd's backward through the same backbone call.
 *   gout            (B,T,2) dLoss/dout, or NULL to use the fused MSE gradient
 *                   gscale * gscale_dev[0] * (out - target)   (gscale = 2/(2*B_global*T) for MSELoss 'mean';
 *                   gscale_dev is an optional device scalar = the upstream dLoss, NULL == 1)
 *   gx              (B,T,2) out (overwritten) when ODPD_F_NEED_DX, else may be NULL
 *   gparams         flat fp32, ACCUMULATED (+=) when ODPD_F_NEED_DW, else may be NULL
 *   workspace       odpd_bwd_workspace_bytes(d)  (may be NULL for a dX-only backward: it then runs the plain serial recurrence)
 * The reduction of per-sequence gradient partials is ordered (no float atomics): results are bit-reproducible
 * run to run for fixed (B,T).
 */
int odpd_backbone_bwd(const OdpdDims *d, const float *x, const float *params, const void *saved, const float *gout,
                      const float *out, const float *target, double gscale, const float *gscale_dev, float *gx,
                      float *gparams, void *workspace, void *stream);

/*
 * Optimiser step on the flat buffers: clip_grad_norm_(max_norm) over `grad` (train_funcs.py:41-42; global L2 norm,
 * coefficient min(1, max_norm/(norm+1e-6)) as torch.nn.utils.clip_grad_norm_) followed by torch.optim.AdamW
 * (project.py:283; decoupled weight decay, bias correction, eps outside the sqrt) — one launch.
 *   step_dev  int64[1] device step counter, incremented by the kernel (keeps the call graph-capturable)
 *   lr_dev    float[1] device learning rate (ReduceLROnPlateau updates it host-side, project.py:288-297)
 *   gnorm_out float[1] or NULL: pre-clip gradient norm
 *   max_norm <= 0 disables clipping (train_funcs.py:41).
 */
int odpd_clip_adamw(float *param, float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, const float *lr_dev,
                    float beta1, float beta2, float eps, float weight_decay, float max_norm, int64_t *step_dev,
                    float *gnorm_out, int zero_grad, void *stream);

/* Optional fusion of the optimiser into the backward (single GPU): arms the NEXT odpd_backbone_bwd call of this host thread (one with
 * ODPD_F_NEED_DW | ODPD_F_OVERWRITE_DW) to run clip_grad_norm_ + AdamW — same arithmetic and arguments as odpd_clip_adamw, on the
 * `gparams` that call produces — inside its gradient-reduction kernel (last CTA to finish), saving one launch per train step; the caller
 * then does NOT call odpd_clip_adamw for that step.  ticket_dev: device int32, zero before first use (the kernel resets it).
 * One-shot, thread-local host state. */
int odpd_fuse_next_bwd_with_adamw(float *param, float *exp_avg, float *exp_avg_sq, const float *lr_dev, float beta1, float beta2, float eps,
                                  float weight_decay, float max_norm, int64_t *step_dev, float *gnorm_out, int32_t *ticket_dev);

/*
 * Data-parallel exchange over NVLink peer memory (ONE exchange per train step, the flat [grad | loss] buffer).
 * The all-reduce is fused into the optimiser kernel and is PUSH based: every rank stores its gradient as 8-byte words
 * {fp32 bits, step tag} straight into a slot of every peer's receive buffer through the NVLink/NVSwitch peer mappings (data and
 * flag travel in one store: no fence, no flag round trip, nothing is read across the link), polls its own LOCAL buffer until every
 * source's words carry this step's tag, sums the sources in rank order (so all replicas compute bit-identical updates), clips
 * and applies AdamW — no NCCL call, no extra launch.  Buffers are symmetric: rank r allocates `odpd_dp_alloc`, exports
 * `odpd_dp_ipc_handle`, the 64-byte handles are all-gathered out of band (torch.distributed), peers map them with `odpd_dp_ipc_open`.
 *   layout of one rank's buffer:  uint2 [2 step parities][world sources][stride]   (stride = n_params + 1 rounded up to 4; element
 *                                 n_params of a source = that rank's loss), zero-initialised by odpd_dp_alloc; world <= 8
 * These are the only entry points that allocate device memory (peer-mappable memory must come from cudaMalloc).
 */
int64_t odpd_dp_buffer_bytes(int64_t n_params);
int odpd_dp_alloc(int64_t bytes, void **out_ptr);
int odpd_dp_free(void *ptr);
int odpd_dp_ipc_handle(void *ptr, unsigned char handle_out[64]);
int odpd_dp_ipc_open(const unsigned char handle[64], void **out_peer_ptr);
int odpd_dp_ipc_close(void *peer_ptr);
/* bufs: HOST array of `world` device pointers (index = rank; bufs[rank] is the caller's own buffer).  grad_local: this rank's flat
 * gradient of the step (`gparams` of odpd_backbone_bwd; NULL = already published by odpd_dp_publish_next_bwd), loss_local its loss
 * (double, from odpd_backbone_fwd).  The slot parity is
 * (step_dev+1)&1, read on the device, so the launch is identical every step and can be replayed from a CUDA graph.  On return
 * (stream order) `param` is updated, loss_out[0] = sum of all ranks' losses, gnorm_out = pre-clip norm of the summed gradient.
 * status_dev[0] != 0 (= 1 + rank of a peer) reports a peer that did not publish within 2 s: the kernel then leaves the parameters
 * and the step counter untouched instead of hanging; the caller must poll status_dev and stop (NativeTrainStep does, and raises). */
/* Optional: arm the NEXT odpd_backbone_bwd call of this host thread (one with ODPD_F_NEED_DW) to publish the gradient it reduces
 * — and the loss `loss_local` — to every rank's receive buffer straight from the epilogue of its gradient reduction, so that the
 * NVLink flight overlaps the launch of the optimiser kernel.  The matching odpd_dp_clip_adamw call then passes grad_local = NULL.
 * One-shot, thread-local host state; an empty batch (B == 0) publishes a zero gradient. */
int odpd_dp_publish_next_bwd(void *const *bufs, int world, int rank, int64_t n, const int64_t *step_dev, const double *loss_local);
int odpd_dp_clip_adamw(float *param, void *const *bufs, int world, int rank, int64_t n, const float *grad_local, const double *loss_local,
                       float *exp_avg, float *exp_avg_sq, const float *lr_dev, float beta1, float beta2, float eps, float weight_decay,
                       float max_norm, int64_t *step_dev, float *gnorm_out, float *loss_out, int *status_dev, void *stream);

/*
 * Evaluation metrics (reference: utils/metrics.py, driven by modules/train_funcs.py:93-105).
 *   odpd_nmse_sums      out[2*s] = sum_n |truth - pred|^2, out[2*s+1] = sum_n |truth|^2 over row s of the (S,N,2) tensors
 *                       (NMSE = mean_s 10 log10(out[2s]/out[2s+1]), utils/metrics.py:42-53).
 *   odpd_dft_magnitude  out[s][g][i] (double, fftshift-ed bin order) = |DFT_nfft of segment g of row s of (a - b)|, b may be NULL;
 *                       segment g = samples g*hop .. g*hop+nfft-1 (zero beyond N: np.fft.fft(x, n) semantics, utils/metrics.py:30),
 *                       optionally with the segment mean removed (`detrend`) and a periodic Hann window (`hann`) — the pieces of
 *                       scipy.signal.welch the reference's power_spectrum uses (utils/metrics.py:158-190).  nfft <= 6144.
 */
int odpd_nmse_sums(const float *pred, const float *truth, int32_t S, int32_t N, double *out, void *stream);
int odpd_dft_magnitude(const float *a, const float *b, int32_t S, int32_t N, int32_t nfft, int32_t nseg, int32_t hop, int32_t hann,
                       int32_t detrend, double *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ODPD_H_ */
