"""CPU restatement of upstream's BubbleForecast sample construction -- TEST INFRASTRUCTURE ONLY (imported by tests/).

Follows bubbleformer/data/dataset.py: __len__ (:62-67), normalize (:69-118), __getitem__ (:120-186).  The upstream
module imports h5py, which is not installed here, so it cannot be executed: PARITY UNPINNED for the indexing logic.
The HDF5 reader it stands on is pinned: tests compare `hdf5_min.read_hdf5` with the byte offsets of
samples/sample_1.hdf5 recorded in the survey and with the committed fixture tests/golden/rollout_sample1_small.npz.
"""
import numpy as np


def dataset_len(traj_lens, time_window, start_time):
    return sum(t - start_time - 2 * time_window + 1 for t in traj_lens)


def norm_terms(arrays, fields, norm):
    diff, div = {}, {}
    for k in fields:
        dl, vl = [], []
        for d in arrays:
            x = np.asarray(d[k])
            if norm == "std":
                dl.append(x.mean()); vl.append(x.std())
            elif norm == "minmax":
                dl.append(x.min()); vl.append(x.max() - x.min())
            elif norm == "tanh":
                dl.append((x.max() + x.min()) / 2.0); vl.append((x.max() - x.min()) / 2.0)
            elif norm == "none":
                dl.append(0.0); vl.append(1.0)
            else:
                raise ValueError(f"Unknown normalization type: {norm}")
        diff[k] = np.mean(dl).item()
        div[k] = np.mean(vl).item() + 1e-8
    return diff, div


def get_item(arrays, idx, input_fields, output_fields, time_window, start_time, diff, div):
    per = [d[input_fields[0]].shape[0] - start_time - 2 * time_window + 1 for d in arrays]
    cum = np.cumsum(per)
    file_idx = np.searchsorted(cum, idx, side="right")
    start = idx + start_time - (cum[file_idx - 1] if file_idx > 0 else 0)
    a, b = slice(start, start + time_window), slice(start + time_window, start + 2 * time_window)
    inp = np.stack([(np.asarray(arrays[file_idx][k][a], dtype=np.float32) - diff[k]) / div[k] for k in input_fields])
    out = np.stack([(np.asarray(arrays[file_idx][k][b], dtype=np.float32) - diff[k]) / div[k] for k in output_fields])
    return inp.astype(np.float32).transpose(1, 0, 2, 3), out.astype(np.float32).transpose(1, 0, 2, 3)
