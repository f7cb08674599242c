"""CPU restatement of upstream's BubbleForecast sample construction -- TEST INFRASTRUCTURE ONLY (imported by tests/).

Follows bubbleformer/data/dataset.py: __len__ (:62-67), normalize (:69-118), __getitem__ (:120-186, including the
nearest-neighbour downsampling of :138-153).  Pinned: oracle/make_data_golden.py executed the UNMODIFIED upstream class on
upstream's two sample files (with a stand-in for the absent h5py that serves the datasets as numpy arrays) and committed
samples, lengths and normalisation constants to tests/golden/dataset.npz; tests/test_data.py checks this restatement
against them.  The HDF5 reader is pinned separately on the raw bytes of samples/sample_1.hdf5.
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


def nearest_index(n_in, factor):
    """Source indices of F.interpolate(mode="nearest") to size n_in // factor (dataset.py:141-147): ATen's
    nearest_neighbor_compute_source_index, src = min(floor(dst * scale), n_in - 1) with scale = n_in / n_out in float32."""
    n_out = n_in // factor
    scale = np.float32(n_in) / np.float32(n_out)
    return np.minimum(np.floor(np.arange(n_out, dtype=np.float32) * scale).astype(np.int64), n_in - 1)


def get_item(arrays, idx, input_fields, output_fields, time_window, start_time, diff, div, downsample_factor=1):
    per = [d[input_fields[0]].shape[0] - start_time - 2 * time_window + 1 for d in arrays]
    cum = np.cumsum(per)
    file_idx = np.searchsorted(cum, idx, side="right")
    start = idx + start_time - (cum[file_idx - 1] if file_idx > 0 else 0)
    a, b = slice(start, start + time_window), slice(start + time_window, start + 2 * time_window)
    def cut(k, sl):
        x = np.asarray(arrays[file_idx][k][sl], dtype=np.float32)
        if downsample_factor > 1:
            x = x[:, nearest_index(x.shape[1], downsample_factor)][:, :, nearest_index(x.shape[2], downsample_factor)]
        return (x - diff[k]) / div[k]
    inp = np.stack([cut(k, a) for k in input_fields])
    out = np.stack([cut(k, b) for k in output_fields])
    return inp.astype(np.float32).transpose(1, 0, 2, 3), out.astype(np.float32).transpose(1, 0, 2, 3)
