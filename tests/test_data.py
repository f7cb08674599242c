"""Input pipeline: dependency-free HDF5 reader pinned on CPU, device window gather against the oracle."""
import os
import struct
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _write_h5_v0(path, arrays):
    """A tiny HDF5 writer (superblock 0, one symbol-table node, contiguous float32 datasets) for the reader test."""
    names = list(arrays)
    heap_data = b"\x00" * 8
    name_off = {}
    for n in names:
        name_off[n] = len(heap_data)
        heap_data += n.encode() + b"\x00"
        heap_data += b"\x00" * (-len(heap_data) % 8)
    heap_data += b"\x00" * 64
    P = {"sb": 0, "root_oh": 96, "btree": 136, "heap": 136 + 24 + 16 * 33 + 8, }
    P["heap_data"] = P["heap"] + 32
    P["snod"] = P["heap_data"] + len(heap_data)
    P["snod"] += -P["snod"] % 8
    oh0 = P["snod"] + 8 + 40 * 32
    oh_size = 16 + (8 + 8 + 8 * 3) + (8 + 24) + (8 + 24)
    oh_size += -oh_size % 8
    data0 = oh0 + oh_size * len(names)
    data0 += -data0 % 2048
    buf = bytearray(data0 + sum(a.nbytes for a in arrays.values()))
    u = lambda off, fmt, *v: struct.pack_into("<" + fmt, buf, off, *v)
    buf[0:8] = b"\x89HDF\r\n\x1a\n"
    buf[8:16] = bytes([0, 0, 0, 0, 0, 8, 8, 0])
    u(16, "HHI", 4, 16, 0)
    u(24, "QQQQ", 0, 0xFFFFFFFFFFFFFFFF, len(buf), 0xFFFFFFFFFFFFFFFF)
    u(56, "QQII", 0, P["root_oh"], 1, 0)
    u(80, "QQ", P["btree"], P["heap"])
    u(P["root_oh"], "BBHII", 1, 0, 1, 1, 24)                 # root object header: one symbol-table message
    u(P["root_oh"] + 16, "HHBBBB", 0x11, 16, 0, 0, 0, 0)
    u(P["root_oh"] + 24, "QQ", P["btree"], P["heap"])
    buf[P["btree"]:P["btree"] + 4] = b"TREE"
    u(P["btree"] + 4, "BBH", 0, 0, 1)
    u(P["btree"] + 8, "QQ", 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF)
    u(P["btree"] + 24, "QQQ", 0, P["snod"], name_off[names[-1]])
    buf[P["heap"]:P["heap"] + 4] = b"HEAP"
    u(P["heap"] + 8, "QQQ", len(heap_data), 0xFFFFFFFFFFFFFFFF, P["heap_data"])
    buf[P["heap_data"]:P["heap_data"] + len(heap_data)] = heap_data
    buf[P["snod"]:P["snod"] + 4] = b"SNOD"
    u(P["snod"] + 4, "BBH", 1, 0, len(names))
    off = data0
    for i, n in enumerate(names):
        a = np.ascontiguousarray(arrays[n], dtype="<f4")
        oh = oh0 + i * oh_size
        u(P["snod"] + 8 + 40 * i, "QQII", name_off[n], oh, 0, 0)
        u(oh, "BBHII", 1, 0, 3, 1, oh_size - 16)
        p = oh + 16
        u(p, "HHBBBB", 0x1, 8 + 8 * a.ndim, 0, 0, 0, 0); u(p + 8, "BBBBI", 1, a.ndim, 0, 0, 0)
        for k, s in enumerate(a.shape):
            u(p + 16 + 8 * k, "Q", s)
        p += 8 + 8 + 8 * 3
        u(p, "HHBBBB", 0x3, 24, 0, 0, 0, 0); u(p + 8, "BBBBI", 0x11, 0x20, 0x1F, 0, 4); u(p + 16, "HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
        p += 8 + 24
        u(p, "HHBBBB", 0x8, 24, 0, 0, 0, 0); u(p + 8, "BB", 3, 1); u(p + 10, "QQ", off, a.nbytes)
        buf[off:off + a.nbytes] = a.tobytes()
        off += a.nbytes
    with open(path, "wb") as f:
        f.write(bytes(buf))


def test_hdf5_reader_roundtrip(tmp_path):
    from bubbleformer_b200.hdf5_min import read_hdf5
    rng = np.random.default_rng(0)
    arrays = {k: rng.standard_normal((7, 8, 12)).astype(np.float32) for k in ("dfun", "temperature", "velx", "vely")}
    p = str(tmp_path / "traj.hdf5")
    _write_h5_v0(p, arrays)
    got = read_hdf5(p)
    assert set(got) == set(arrays)
    for k in arrays:
        assert got[k].dtype == np.float32 and np.array_equal(got[k], arrays[k])


def test_hdf5_reader_matches_committed_sample_frames():
    """tests/golden/rollout_sample1_small.npz holds frames of upstream's samples/sample_1.hdf5 (written with the
    survey's byte offsets); when the reference checkout is present the reader must reproduce them."""
    from bubbleformer_b200.hdf5_min import read_hdf5
    ref = "/root/reference/samples/sample_1.hdf5"
    if not os.path.exists(ref):
        pytest.skip("reference checkout not mounted (GPU box)")
    d = read_hdf5(ref)
    assert {k: v.shape for k, v in d.items()} == {k: (50, 64, 64) for k in ("dfun", "temperature", "velx", "vely")}
    with open(ref, "rb") as f:
        buf = f.read()
    for k, off in {"dfun": 2048, "temperature": 821248, "velx": 1640448, "vely": 2461696}.items():
        assert np.array_equal(d[k], np.frombuffer(buf, dtype="<f4", count=50 * 64 * 64, offset=off).reshape(50, 64, 64))


def test_data_oracle_index_arithmetic():
    from oracle import data_oracle as O
    arrays = [{"dfun": np.arange(30 * 4, dtype=np.float32).reshape(30, 2, 2)}, {"dfun": 1000 + np.arange(26 * 4, dtype=np.float32).reshape(26, 2, 2)}]
    T, st = 3, 5
    assert O.dataset_len([30, 26], T, st) == (30 - 5 - 6 + 1) + (26 - 5 - 6 + 1)
    diff, div = O.norm_terms(arrays, ["dfun"], "none")
    inp, out = O.get_item(arrays, 0, ["dfun"], ["dfun"], T, st, diff, div)
    assert inp.shape == (3, 1, 2, 2) and inp[0, 0, 0, 0] == 5 * 4 and out[0, 0, 0, 0] == 8 * 4
    inp, out = O.get_item(arrays, 20, ["dfun"], ["dfun"], T, st, diff, div)        # first sample of the second file
    assert inp[0, 0, 0, 0] == 1000 + 5 * 4


@pytest.mark.gpu
@pytest.mark.parametrize("norm", ["none", "std", "minmax", "tanh"])
def test_device_windows_match_oracle(norm, tmp_path):
    import json
    import torch
    from bubbleformer_b200.data import DeviceForecastWindows
    from oracle import data_oracle as O
    rng = np.random.default_rng(1)
    fields = ["dfun", "temperature", "velx", "vely"]
    arrays = [{k: (rng.standard_normal((n, 16, 24)) * (1 + i)).astype(np.float32) for i, k in enumerate(fields)} for n in (40, 33)]
    files = []
    for j, a in enumerate(arrays):
        p = str(tmp_path / f"t{j}.hdf5")
        _write_h5_v0(p, a)
        with open(p.replace(".hdf5", ".json"), "w") as f:
            json.dump(dict(inv_reynolds=0.1 + j, cpgas=1.0, mugas=2.0, rhogas=3.0, thcogas=4.0, stefan=5.0, prandtl=6.0,
                           heater=dict(nucWaitTime=7.0, wallTemp=90.0 + j)), f)
        files.append(p)
    T, st = 5, 3
    ds = DeviceForecastWindows(files, input_fields=fields, output_fields=["temperature", "dfun"], norm=norm, time_window=T,
                               start_time=st, return_fluid_params=True)
    diff, div = ds.normalize()
    rd, rv = O.norm_terms(arrays, ds.fields, norm)
    assert all(abs(diff[k] - rd[k]) < 1e-6 and abs(div[k] - rv[k]) < 1e-6 * rv[k] for k in ds.fields)
    assert len(ds) == O.dataset_len([40, 33], T, st)
    idx = [0, 1, len(ds) - 1, 27, 28, 13]
    inp, tgt, cond = ds.batch(idx)
    assert inp.shape == (6, T, 4, 16, 24) and tgt.shape == (6, T, 2, 16, 24) and cond.shape == (6, 9)
    for b, i in enumerate(idx):
        ri, ro = O.get_item(arrays, i, fields, ["temperature", "dfun"], T, st, rd, rv)
        assert float((inp[b].cpu() - torch.from_numpy(ri)).abs().max()) < 1e-5 * max(1.0, float(np.abs(ri).max()))
        assert float((tgt[b].cpu() - torch.from_numpy(ro)).abs().max()) < 1e-5 * max(1.0, float(np.abs(ro).max()))
    assert float(cond[2, 0]) == pytest.approx(1.1) and float(cond[0, 8]) == pytest.approx(90.0)


def _golden_dataset():
    import json
    z = np.load(os.path.join(ROOT, "tests", "golden", "dataset.npz"))
    return z, json.loads(str(z["meta"]))


def _sample_arrays():
    """The two upstream sample trajectories: from the reference checkout when mounted, else rebuilt from the fixture of
    case 0 is not possible (it holds windows only) -> skip."""
    from bubbleformer_b200.hdf5_min import read_hdf5
    files = [f"/root/reference/samples/sample_{i}.hdf5" for i in (1, 2)]
    if not all(os.path.exists(f) for f in files):
        pytest.skip("reference checkout not mounted (GPU box)")
    return files, [read_hdf5(f) for f in files]


def test_data_oracle_matches_reference_dataset_fixture():
    """oracle/data_oracle.py against samples produced by the UNMODIFIED upstream BubbleForecast (oracle/make_data_golden.py):
    length, normalisation constants, window index arithmetic across the file boundary, nearest downsampling."""
    from oracle import data_oracle as O
    z, meta = _golden_dataset()
    files, arrays = _sample_arrays()
    for ci, m in enumerate(meta):
        fields = sorted(set(m["input_fields"] + m["output_fields"]))
        assert O.dataset_len([50, 50], m["time_window"], m["start_time"]) == m["length"]
        diff, div = O.norm_terms(arrays, fields, m["norm"])
        for k in fields:
            assert abs(diff[k] - m["diff"][k]) <= 1e-6 * max(1.0, abs(m["diff"][k])), (ci, k)
            assert abs(div[k] - m["div"][k]) <= 1e-6 * max(1.0, abs(m["div"][k])), (ci, k)
        for i in m["indices"]:
            inp, tgt = O.get_item(arrays, i, m["input_fields"], m["output_fields"], m["time_window"], m["start_time"],
                                  m["diff"], m["div"], m["downsample_factor"])
            assert inp.shape == z[f"c{ci}_i{i}_inp"].shape
            assert np.allclose(inp, z[f"c{ci}_i{i}_inp"], rtol=1e-6, atol=1e-6), (ci, i)
            assert np.allclose(tgt, z[f"c{ci}_i{i}_tgt"], rtol=1e-6, atol=1e-6), (ci, i)


@pytest.mark.gpu
def test_device_windows_downsample_and_fields_match_pinned_oracle(tmp_path):
    """The device pipeline (hdf5_min reader + resident frames + bf_window_gather, downsample_factor included) on the
    configurations of tests/golden/dataset.npz against oracle/data_oracle.py -- which the CPU test above pins to the
    samples the UNMODIFIED upstream BubbleForecast produced.  (The GPU box has no upstream sample files: the
    trajectories are seeded arrays written through the test's HDF5 writer.)"""
    import torch
    from bubbleformer_b200.data import DeviceForecastWindows
    from oracle import data_oracle as O
    _, meta = _golden_dataset()
    rng = np.random.default_rng(5)
    arrays = [{k: (rng.standard_normal((50, 64, 64)) * (1 + j) + j).astype(np.float32)
               for j, k in enumerate(("dfun", "temperature", "velx", "vely"))} for _ in range(2)]
    files = []
    for j, a in enumerate(arrays):
        p = str(tmp_path / f"s{j}.hdf5")
        _write_h5_v0(p, a)
        files.append(p)
    for m in meta:
        ds = DeviceForecastWindows(files, input_fields=m["input_fields"], output_fields=m["output_fields"], norm=m["norm"],
                                   time_window=m["time_window"], start_time=m["start_time"],
                                   downsample_factor=m["downsample_factor"])
        ds.normalize()
        assert len(ds) == m["length"]
        rd, rv = O.norm_terms(arrays, ds.fields, m["norm"])
        inp, tgt = ds.batch(m["indices"])
        f = m["downsample_factor"]
        assert inp.shape == (len(m["indices"]), m["time_window"], len(m["input_fields"]), 64 // f, 64 // f)
        for b, i in enumerate(m["indices"]):
            ri, ro = O.get_item(arrays, i, m["input_fields"], m["output_fields"], m["time_window"], m["start_time"], rd, rv, f)
            assert tuple(inp[b].shape) == ri.shape and tuple(tgt[b].shape) == ro.shape
            assert float((inp[b].cpu() - torch.from_numpy(ri)).abs().max()) < 1e-5 * max(1.0, float(np.abs(ri).max()))
            assert float((tgt[b].cpu() - torch.from_numpy(ro)).abs().max()) < 1e-5 * max(1.0, float(np.abs(ro).max()))


@pytest.mark.parametrize("norm", ["none", "std", "minmax", "tanh"])
def test_device_windows_host_side_pieces(norm):
    """The host-side pieces of DeviceForecastWindows that need no GPU: the per-file normalisation terms (taken at
    construction, so the host copies of the fields can be dropped) average to the oracle's constants, and the
    channel-by-channel upload lays the trajectories out as (sum frames, C, H, W) like the stacked copy it replaced."""
    import torch
    from bubbleformer_b200.data import DeviceForecastWindows as D
    from oracle import data_oracle as O
    rng = np.random.default_rng(3)
    fields = ["dfun", "temperature", "velx", "vely"]
    arrays = [{k: (rng.standard_normal((n, 6, 8)) * (1 + i) + i).astype(np.float32) for i, k in enumerate(fields)} for n in (7, 5)]
    terms = [{k: D._field_terms(a[k], norm) for k in fields} for a in arrays]
    rd, rv = O.norm_terms(arrays, fields, norm)
    for k in fields:
        assert abs(np.mean([t[k][0] for t in terms]).item() - rd[k]) < 1e-6
        assert abs(np.mean([t[k][1] for t in terms]).item() + 1e-8 - rv[k]) < 1e-6 * rv[k]
    frames = D._upload(arrays, fields, torch.device("cpu"))
    ref = np.concatenate([np.stack([a[k] for k in fields], axis=1) for a in arrays], axis=0)
    assert frames.shape == (12, 4, 6, 8) and np.array_equal(frames.numpy(), ref)
