// cells.h — internal (non-ABI) interfaces between api.cu and the per-family kernel files.
#pragma once
#include "common.cuh"

namespace odpd {

struct GruArgs {
    int B, T, H;
    const float *x, *target, *params, *gout, *out_in, *gscale_dev;
    float *out, *gx, *partials;
    double *loss;
    float *saved;
    float loss_scale, gscale;
    int save, need_dx;
    // delta cells / DVRJANET
    float thx, thh;
    int64_t *stats;
    int K, cell;
    // time-chunked ("speculative") execution of the contractive GRU-family recurrences (gru_family.cu):
    //   C      chunks per sequence (1 = plain serial), Lc = steps per chunk (multiple of 32), Wu = warm-up steps
    //   mode   0 = run (serial when C==1, else one CTA per (sequence, chunk)), 2 = verify every chunk boundary and re-run the
    //          sequences that failed serially
    //   sc_guess/sc_end [B*C][HP]: state a chunk started from after its warm-up / state it ended with;  sc_loss [B*C]: per-chunk
    //   squared error;  sc_fail: number of sequences that needed the serial re-run (diagnostic)
    // storage of the IQ tensors: bf16 pairs instead of fp32 pairs; frame starts into a raw (N,2) stream instead of (B,T,2) frames
    int x_bf16, target_bf16;
    const int *x_starts, *target_starts;
    int C, Lc, Wu, mode;
    float *sc_guess, *sc_end, *sc_loss;
    int *sc_fail;
    float tol;
    int tchunks_req, twarm_req, twarm_default;
    int wt_blocks;   // split backward: 32-step blocks per CTA of the weights kernel
    float *vd_contrib;   // VDLSTM backward: per-(step, window tap) dL/dx contributions [B][T][4] float2 (lstm.cu)
    float *gbuf;     // split backward (gru_family.cu): per-step gate gradients [B][T][4*HP+4], written by the chain kernel, read by the weights kernel
};

// gru_family.cu : GRU / DGRU / QGRU / QGRU_AMP1
int64_t gru_family_nparams(int cell, int H);
int64_t gru_family_saved_floats(int cell, int B, int T, int H, bool save, int tchunks_req);
int64_t gru_family_workspace_floats(int cell, int B, int T, int H, int tchunks_req);
int gru_family_run(int cell, const GruArgs &a, int dir, bool dw, cudaStream_t st, int *rows_out);
int gru_family_plan(int cell, int B, int T, int H, int tchunks_req, int twarm_req, int dir, bool dw, bool save, int out[4]);

#define ODPD_HAVE_DELTA 1
#define ODPD_HAVE_JANET 1
#define ODPD_HAVE_GMP 1
// lstm.cu
int64_t lstm_saved_floats(int B, int T, int H, bool save, int tchunks_req);
int64_t lstm_workspace_floats(int B, int T, int H, int64_t P, int tchunks_req, bool vd);
int64_t lstm_nparams(int H, bool vd);
int lstm_run(const GruArgs &a, int dir, bool dw, cudaStream_t st, int *info);   // dir +2 = plan only; info[0..3] see chunking.cuh
// delta.cu : DELTAGRU / TRES
int64_t delta_saved_floats(int cell, int B, int T, int H);
int delta_run(const GruArgs &a, int dir, bool dw, cudaStream_t st);
// janet.cu : PGJANET / DVRJANET
int64_t janet_saved_floats(int cell, int B, int T, int H, bool save, int tchunks_req);
int64_t janet_workspace_floats(int cell, int B, int H, int64_t P, int tchunks_req);
int janet_run(const GruArgs &a, int dir, bool dw, cudaStream_t st, int *info);
// qgru_qat.cu : fake-quantised GRU (QAT)
int64_t qat_nparams(int H);
int64_t qat_saved_floats(int B, int T, int H);
int qat_run(const GruArgs &a, int dir, bool dw, cudaStream_t st);
// gmp.cu
int gmp_run(const GruArgs &a, int dir, bool dw, cudaStream_t st);

// wide.cu : GRU / LSTM / DGRU / QGRU / QGRU_AMP1 with hidden_size 33..64 and/or num_layers > 1 (layered, unchunked)
bool wide_supported(int cell);
int64_t wide_nparams(int cell, int H, int layers);
int64_t wide_saved_floats(int cell, int B, int T, int H, int layers, bool save);
int64_t wide_workspace_floats(int cell, int B, int T, int H, int layers);
int wide_run(int cell, const GruArgs &a, int layers, int dir, bool dw, cudaStream_t st, int *rows_out);

// rvtdcnn.cu : RVTDCNN (feed-forward over a 4-sample wrap-around window; nothing saved, the backward recomputes)
int64_t rvtdcnn_nparams(int H);
int64_t rvtdcnn_workspace_floats(int B, int T, int H);
int rvtdcnn_run(const GruArgs &a, int dir, bool dw, cudaStream_t st, int *rows_out);

// bojanet.cu : BOJANET (FIR front end + vector demodulator + f/g recurrence + phase rotation)
int64_t bojanet_nparams(int H);
int64_t bojanet_saved_floats(int B, int T, int H);
int64_t bojanet_workspace_floats(int B, int T, int H);
int bojanet_run(const GruArgs &a, int dir, bool dw, cudaStream_t st, int *rows_out);

// tcnn.cu : TCNN / NeuralTX (dilated depthwise temporal-convolution stack)
int64_t tcnn_nparams(int cell, int C);
int64_t tcnn_saved_floats(int B, int T, int C);
int64_t tcnn_workspace_floats(int cell, int B, int T, int C);
int tcnn_run(int cell, const GruArgs &a, int dir, bool dw, cudaStream_t st, int *rows_out);

// apnrru.cu : APNRRU (phase-normalised recurrent unit)
int64_t apnrru_nparams(int H);
int64_t apnrru_saved_floats(int B, int T, int H);
int64_t apnrru_workspace_floats(int B, int T, int H);
int apnrru_run(const GruArgs &a, int dir, bool dw, cudaStream_t st, int *rows_out);

// mcldnn.cu : MCLDNN (linear convolutional front end composed in parameter space + LSTM(8) + two linear layers)
int64_t mcldnn_nparams(int C);
int64_t mcldnn_saved_floats(int B, int T, int C);
int64_t mcldnn_workspace_floats(int B, int T, int C);
int mcldnn_run(const GruArgs &a, int dir, bool dw, cudaStream_t st, int *rows_out);

// deltajanet.cu : DeltaJANET (JANET cell computed incrementally; the reference fixes its thresholds at zero)
int64_t deltajanet_nparams(int H);
int64_t deltajanet_saved_floats(int B, int T, int H);
int64_t deltajanet_workspace_floats(int B, int T, int H);
int deltajanet_run(const GruArgs &a, int dir, bool dw, cudaStream_t st, int *rows_out);

// tres_qat.cu : fake-quantised TRes-DeltaGRU (QAT)
int64_t tresq_nparams(int H);
int64_t tresq_saved_floats(int B, int T);
int64_t tresq_workspace_floats(int B, int T, int H);
int tresq_run(const GruArgs &a, int dir, bool dw, cudaStream_t st, int *rows_out);

// remaining families (lstm.cu, delta.cu, janet.cu, gmp.cu) behind one dispatcher in others.cu
int64_t other_nparams(int cell, int H, int K);
int64_t other_saved_bytes(const OdpdDims *d);
int64_t other_workspace_floats(const OdpdDims *d);
int other_plan(const OdpdDims *d, int backward, int out[4]);
int other_fwd(const OdpdDims *d, const float *x, const float *target, const float *params, float *out, double *loss, double loss_scale,
              void *saved, int64_t *stats, cudaStream_t st);
int other_bwd(const OdpdDims *d, const float *x, const float *params, const void *saved, const float *gout, const float *out,
              const float *target, double gscale, const float *gscale_dev, float *gx, float *partials, cudaStream_t st, int *rows_out);

int reduce_partials(const float *part, int nrows, int64_t P, float *g, int overwrite, cudaStream_t st, const DpPushArgs *push = nullptr,
                    const AdamFuseArgs *adam = nullptr);
// one-shot data-parallel publish context armed by odpd_dp_publish_next_bwd (dp.cu) and consumed by the next weight-gradient reduction
bool dp_take_armed_push(DpPushArgs &out);

}  // namespace odpd
