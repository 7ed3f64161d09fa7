// others.cu — dispatcher for the non-GRU families.
#include "cells.h"

namespace odpd {

int64_t other_nparams(int cell, int H, int K) {
    switch (cell) {
    case ODPD_CELL_LSTM: return (int64_t)4 * H * 2 + 4 * H * H + 8 * H + 2 * H + 2;
    case ODPD_CELL_DELTAGRU: return (int64_t)3 * H * 6 + 3 * H * H + 6 * H + 2 * H + 2;
    case ODPD_CELL_TRES: return (int64_t)3 * H * 6 + 3 * H * H + 2 * H + 18 + 6;
    case ODPD_CELL_PGJANET: return (int64_t)3 * (H * (H + 1) + H) + 2 * (2 * H * H + H) + 2 * H + 2;
    case ODPD_CELL_DVRJANET: return (int64_t)K + 3 * H * H + 2 * H + H + 2 * (2 * H * H + H) + 2 * (H + 1);
    case ODPD_CELL_GMP: return 495;
    }
    return -1;
}

int64_t other_saved_bytes(const OdpdDims *d) {
    set_error("cell %d: not implemented yet", d->cell);
    return -1;
}
int other_fwd(const OdpdDims *d, const float *, const float *, const float *, float *, double *, double, void *, int64_t *, cudaStream_t) {
    set_error("cell %d: forward not implemented yet", d->cell);
    return -3;
}
int other_bwd(const OdpdDims *d, const float *, const float *, const void *, const float *, const float *, const float *, double,
              const float *, float *, float *, cudaStream_t) {
    set_error("cell %d: backward not implemented yet", d->cell);
    return -3;
}

}  // namespace odpd
