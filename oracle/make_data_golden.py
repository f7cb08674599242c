"""Generate tests/golden/dataset.npz by executing the UNMODIFIED reference `BubbleForecast` -- TEST INFRASTRUCTURE ONLY.

The reference dataset (bubbleformer/data/dataset.py) imports h5py, which is not installed in this image.  This script
registers a stand-in `h5py` module whose `File(name, "r")` returns the datasets of the file as numpy arrays read by
bubbleformer_b200/hdf5_min.py (itself pinned on the raw bytes of upstream's samples/sample_1.hdf5,
tests/test_data.py).  Everything else -- __len__, normalize, the index arithmetic of __getitem__, the nearest-neighbour
downsampling through F.interpolate, the field order, the fluid-parameter vector -- is the live reference code.
Runs only where the upstream checkout is mounted (/root/reference).      python oracle/make_data_golden.py
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("BUBBLEFORMER_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

CASES = [
    # (input_fields, output_fields, norm, downsample_factor, time_window, start_time, indices)
    (["dfun", "temperature", "velx", "vely"], ["dfun", "temperature", "velx", "vely"], "none", 1, 5, 5, [0, 36]),
    (["dfun", "temperature", "velx", "vely"], ["temperature", "velx"], "std", 2, 5, 5, [3, 40]),
    (["temperature", "velx", "vely"], ["dfun"], "minmax", 4, 10, 5, [0, 25, 26, 51]),
    (["dfun"], ["temperature", "velx", "vely"], "tanh", 2, 10, 3, [7, 30]),
]


def main():
    from bubbleformer_b200.hdf5_min import read_hdf5
    shim = types.ModuleType("h5py")
    shim.File = lambda name, mode="r": read_hdf5(name)
    sys.modules["h5py"] = shim
    spec = importlib.util.spec_from_file_location("_ref_dataset", os.path.join(REF, "bubbleformer/data/dataset.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    files = [os.path.join(REF, "samples/sample_1.hdf5"), os.path.join(REF, "samples/sample_2.hdf5")]
    out = {}
    meta = []
    for ci, (fi, fo, norm, ds_f, tw, st, idx) in enumerate(CASES):
        ds = mod.BubbleForecast(filenames=files, input_fields=fi, output_fields=fo, norm=norm, downsample_factor=ds_f,
                                time_window=tw, start_time=st)
        diff, div = ds.normalize()
        meta.append(dict(input_fields=fi, output_fields=fo, norm=norm, downsample_factor=ds_f, time_window=tw,
                         start_time=st, indices=idx, length=len(ds), diff={k: float(v) for k, v in diff.items()},
                         div={k: float(v) for k, v in div.items()}))
        for i in idx:
            inp, tgt = ds[i]
            out[f"c{ci}_i{i}_inp"] = inp.numpy().astype(np.float32)
            out[f"c{ci}_i{i}_tgt"] = tgt.numpy().astype(np.float32)
    path = os.path.join(ROOT, "tests", "golden", "dataset.npz")
    np.savez_compressed(path, meta=json.dumps(meta), **out)
    print(path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
