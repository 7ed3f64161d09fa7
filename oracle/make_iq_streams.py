"""Build tests/golden/iq_streams.npz — the measured IQ training streams of the reference's shipped datasets, as the float32 tensors
`IQFrameDataset` feeds `net_train` (TEST/BENCH INFRASTRUCTURE; run in the authoring container, where /root/reference exists).

    python oracle/make_iq_streams.py

Loading goes through the UNMODIFIED reference (`modules/data_collector.py:16-110 load_dataset`), the casts are the reference's:
  * frames are `torch.Tensor(np.ndarray[float64])` (data_collector.py:236-237) -> float32(x), float32(y) for train_pa;
  * for train_dpd the target is computed in float64 first (`project.py:221-225`: y = target_gain * X) and then cast, so the
    fixture also stores `<ds>.dpd_target = float32(gain * x_f64)` (NOT gain * float32(x)) and the gain
    (`utils/util.py:27-35 set_target_gain`).
bench.py and the -m gpu tests read only the .npz (the GPU box has no /root/reference).  With stride-1 framing (data_collector.py:240-247)
frame k of length T is rows [k, k+T) of a stream, so the stream plus start indices reproduces every frame bit for bit."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DATASETS = ("APA_200MHz", "APA_200MHz_b", "DPA_100MHz", "DPA_200MHz")     # BASELINE.json configs[1..4] / configs[0]


def main():
    os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
    sys.path.insert(0, REF)
    from modules.data_collector import load_dataset
    from utils.util import set_target_gain
    out = {}
    for ds in DATASETS:
        X_train, y_train, X_val, y_val, X_test, y_test = load_dataset(dataset_name=ds)
        X_train, y_train = np.asarray(X_train, dtype=np.float64), np.asarray(y_train, dtype=np.float64)
        gain = float(set_target_gain(X_train, y_train))
        out[ds + ".x"] = X_train.astype(np.float32)
        out[ds + ".y"] = y_train.astype(np.float32)
        out[ds + ".dpd_target"] = (gain * X_train).astype(np.float32)
        out[ds + ".gain"] = np.float64(gain)
        print(ds, X_train.shape, "gain", gain, "max|x|", np.abs(X_train).max(), "min amp", np.sqrt((X_train ** 2).sum(1)).min())
    # the validation segment of the headline dataset (net_eval shape: one nperseg=19662 segment)
    _, _, X_val, y_val, _, _ = load_dataset(dataset_name="APA_200MHz")
    out["APA_200MHz.x_val"] = np.asarray(X_val, dtype=np.float64).astype(np.float32)
    out["APA_200MHz.y_val"] = np.asarray(y_val, dtype=np.float64).astype(np.float32)
    path = os.path.join(ROOT, "tests", "golden", "iq_streams.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
