// others.cu — dispatcher for the non-GRU families.
#include "cells.h"

namespace odpd {

int64_t other_nparams(int cell, int H, int K) {
    switch (cell) {
    case ODPD_CELL_LSTM: return lstm_nparams(H, false);
    case ODPD_CELL_VDLSTM: return lstm_nparams(H, true);
    case ODPD_CELL_DELTAGRU: return (int64_t)3 * H * 6 + 3 * H * H + 6 * H + 2 * H + 2;
    case ODPD_CELL_TRES: return (int64_t)3 * H * 6 + 3 * H * H + 2 * H + 18 + 6;
    case ODPD_CELL_PGJANET: return (int64_t)3 * (H * (H + 1) + H) + 2 * (2 * H * H + H) + 2 * H + 2;
    case ODPD_CELL_DVRJANET: return (int64_t)K + 3 * H * H + 2 * H + H + 2 * (2 * H * H + H) + 2 * (H + 1);
    case ODPD_CELL_GMP: return 495;
    case ODPD_CELL_QGRU_QAT: case ODPD_CELL_QGRU_AMP1_QAT: return qat_nparams(H);
    }
    return -1;
}

static bool implemented(int cell) {
    switch (cell) {
    case ODPD_CELL_LSTM: case ODPD_CELL_VDLSTM: case ODPD_CELL_QGRU_QAT: case ODPD_CELL_QGRU_AMP1_QAT: return true;
#ifdef ODPD_HAVE_DELTA
    case ODPD_CELL_DELTAGRU: case ODPD_CELL_TRES: return true;
#endif
#ifdef ODPD_HAVE_JANET
    case ODPD_CELL_PGJANET: case ODPD_CELL_DVRJANET: return true;
#endif
#ifdef ODPD_HAVE_GMP
    case ODPD_CELL_GMP: return true;
#endif
    }
    return false;
}

static bool chunkable(int cell) { return cell == ODPD_CELL_LSTM || cell == ODPD_CELL_VDLSTM || cell == ODPD_CELL_PGJANET || cell == ODPD_CELL_DVRJANET; }

// bytes of `saved` for these dims: activations (with ODPD_F_SAVE) + the chunk scratch of the chunkable cells
int64_t other_saved_bytes(const OdpdDims *d) {
    int64_t n = -1;
    const bool save = (d->flags & ODPD_F_SAVE) != 0;
    switch (d->cell) {
    case ODPD_CELL_LSTM: case ODPD_CELL_VDLSTM: n = lstm_saved_floats(d->B, d->T, d->H, save, d->tchunks); break;
    case ODPD_CELL_QGRU_QAT: case ODPD_CELL_QGRU_AMP1_QAT: n = qat_saved_floats(d->B, d->T, d->H); break;
#ifdef ODPD_HAVE_DELTA
    case ODPD_CELL_DELTAGRU: case ODPD_CELL_TRES: n = delta_saved_floats(d->cell, d->B, d->T, d->H); break;
#endif
#ifdef ODPD_HAVE_JANET
    case ODPD_CELL_PGJANET: case ODPD_CELL_DVRJANET: n = janet_saved_floats(d->cell, d->B, d->T, d->H, save, d->tchunks); break;
#endif
#ifdef ODPD_HAVE_GMP
    case ODPD_CELL_GMP: n = 4; break;
#endif
    }
    if (n < 0) { set_error("cell %d (H=%d): not available in this build", d->cell, d->H); return -1; }
    if (!save && !chunkable(d->cell)) return 0;
    return 4 * n;
}

int64_t other_workspace_floats(const OdpdDims *d) {
    const int64_t P = other_nparams(d->cell, d->H, d->K);
    switch (d->cell) {
    case ODPD_CELL_LSTM: case ODPD_CELL_VDLSTM: return lstm_workspace_floats(d->B, d->T, d->H, P, d->tchunks, d->cell == ODPD_CELL_VDLSTM);
#ifdef ODPD_HAVE_JANET
    case ODPD_CELL_PGJANET: case ODPD_CELL_DVRJANET: return janet_workspace_floats(d->cell, d->B, d->H, P, d->tchunks);
#endif
    }
    return (int64_t)(d->B > 0 ? d->B : 1) * P;
}

static int run(const OdpdDims *d, const GruArgs &a, int dir, bool dw, cudaStream_t st, int *info) {
    if (!implemented(d->cell)) { set_error("cell %d: not implemented in this build", d->cell); return -3; }
    if (info) { info[0] = 1; info[1] = a.T; info[2] = 0; info[3] = -1; }
    if (dir >= 2 && !chunkable(d->cell)) return 0;
    switch (d->cell) {
    case ODPD_CELL_LSTM: case ODPD_CELL_VDLSTM: return lstm_run(a, dir, dw, st, info);
    case ODPD_CELL_QGRU_QAT: case ODPD_CELL_QGRU_AMP1_QAT: return qat_run(a, dir, dw, st);
#ifdef ODPD_HAVE_DELTA
    case ODPD_CELL_DELTAGRU: case ODPD_CELL_TRES: return delta_run(a, dir, dw, st);
#endif
#ifdef ODPD_HAVE_JANET
    case ODPD_CELL_PGJANET: case ODPD_CELL_DVRJANET: return janet_run(a, dir, dw, st, info);
#endif
#ifdef ODPD_HAVE_GMP
    case ODPD_CELL_GMP: return gmp_run(a, dir, dw, st);
#endif
    }
    return -3;
}

int other_fwd(const OdpdDims *d, const float *x, const float *target, const float *params, float *out, double *loss, double loss_scale,
              void *saved, int64_t *stats, cudaStream_t st) {
    GruArgs a{};
    a.B = d->B; a.T = d->T; a.H = d->H; a.K = d->K; a.cell = d->cell; a.thx = d->thx; a.thh = d->thh; a.stats = stats;
    a.x = x; a.target = target; a.params = params; a.out = out; a.loss = loss; a.loss_scale = (float)loss_scale;
    a.saved = (float *)saved; a.save = (d->flags & ODPD_F_SAVE) != 0;
    a.tchunks_req = d->tchunks; a.twarm_req = d->twarm;
    a.x_bf16 = (d->flags & ODPD_F_X_BF16) != 0; a.target_bf16 = (d->flags & ODPD_F_TARGET_BF16) != 0;
    a.x_starts = d->x_starts; a.target_starts = d->target_starts;
    return run(d, a, 0, false, st, nullptr);
}

int other_bwd(const OdpdDims *d, const float *x, const float *params, const void *saved, const float *gout, const float *out,
              const float *target, double gscale, const float *gscale_dev, float *gx, float *partials, cudaStream_t st, int *rows_out) {
    GruArgs a{};
    a.B = d->B; a.T = d->T; a.H = d->H; a.K = d->K; a.cell = d->cell; a.thx = d->thx; a.thh = d->thh;
    a.x = x; a.params = params; a.saved = (float *)saved; a.gout = gout; a.out_in = out; a.target = target;
    a.gscale = (float)gscale; a.gscale_dev = gscale_dev; a.gx = gx; a.partials = partials;
    a.need_dx = (d->flags & ODPD_F_NEED_DX) != 0;
    a.tchunks_req = d->tchunks; a.twarm_req = d->twarm;
    a.x_bf16 = (d->flags & ODPD_F_X_BF16) != 0; a.target_bf16 = (d->flags & ODPD_F_TARGET_BF16) != 0;
    a.x_starts = d->x_starts; a.target_starts = d->target_starts;
    int info[4] = {1, 0, 0, -1};
    const int rc = run(d, a, 1, (d->flags & ODPD_F_NEED_DW) != 0, st, info);
    if (rows_out) *rows_out = d->B * info[0];
    return rc;
}

int other_plan(const OdpdDims *d, int backward, int out[4]) {
    GruArgs a{};
    a.B = d->B; a.T = d->T; a.H = d->H; a.K = d->K; a.cell = d->cell; a.save = (d->flags & ODPD_F_SAVE) != 0;
    a.tchunks_req = d->tchunks; a.twarm_req = d->twarm;
    return run(d, a, 2 + (backward ? 1 : 0), (d->flags & ODPD_F_NEED_DW) != 0, nullptr, out);
}

}  // namespace odpd
