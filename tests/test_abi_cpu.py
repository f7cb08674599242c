"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol
include/bubbleformer_b200.h declares; ctypes structs match the C layouts; registry / error behaviour mirrors
upstream bubbleformer/models/_api.py; state_dict names and shapes match the reference inventory."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "bubbleformer_b200.h")


def test_library_exports_every_declared_symbol():
    from bubbleformer_b200 import _lib
    src = open(HEADER).read()
    declared = set(re.findall(r"BF_API\s+[\w\s\*]+?\b(bf_\w+)\s*\(", src))
    assert declared and declared == set(_lib.EXPORTS)
    for name in declared:
        assert getattr(_lib.lib, name) is not None
    assert _lib.lib.bf_version() == 100
    assert _lib.lib.bf_launch_count() == 0


def test_ctypes_struct_layouts_match_header(tmp_path):
    from bubbleformer_b200 import _lib
    prog = tmp_path / "sz.cpp"
    prog.write_text(f'#include "{HEADER}"\n#include <cstdio>\nint main(){{printf("%zu %zu %zu %zu %zu\\n",'
                    'sizeof(bf_gemm_args),sizeof(bf_attn_args),sizeof(bf_inorm_apply_args),sizeof(bf_inorm_bwd_args),'
                    'sizeof(bf_inorm_bwd_params_args));}\n')
    exe = tmp_path / "sz"
    subprocess.run(["g++", str(prog), "-o", str(exe)], check=True)
    sizes = [int(t) for t in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [ctypes.sizeof(c) for c in (_lib.GemmArgs, _lib.AttnArgs, _lib.InormApplyArgs, _lib.InormBwdArgs,
                                                 _lib.InormBwdParamsArgs)]


def test_argument_validation_without_gpu():
    """Bad arguments are rejected on the host before any launch, with a readable message."""
    from bubbleformer_b200 import _lib
    a = _lib.GemmArgs()
    assert _lib.lib.bf_gemm(ctypes.byref(a), None) == 1
    assert b"empty problem" in _lib.lib.bf_last_error()
    with pytest.raises(_lib.BubbleformerB200Error):
        _lib.check(1, "bf_gemm")


def test_registry_mirrors_upstream_api():
    from bubbleformer_b200.models import MODELS, get_model, list_models, register_model
    assert {"avit", "filmavit"} <= set(MODELS)
    assert list_models() == sorted(MODELS)
    with pytest.raises(KeyError):
        get_model("vit")                      # upstream's dangling vit_small.yaml raises the same way
    with pytest.raises(ValueError):
        register_model("avit")(object)
    with pytest.raises(AssertionError):
        get_model("avit", patch_size=12)      # non-power-of-two patch (upstream patching.py:24)
    m = get_model("FiLMAViT", input_fields=4, output_fields=4, time_window=5, patch_size=16, embed_dim=384,
                  processor_blocks=2, num_heads=6, drop_path=0.2, attn_scale=True, feat_scale=True, num_fluid_params=9)
    assert type(m).__name__ == "FiLMConditionedAViT"


@pytest.mark.parametrize("film", [True, False])
def test_state_dict_inventory_matches_reference(film):
    from bubbleformer_b200 import get_model
    from oracle.param_init import param_shapes
    cfg = dict(input_fields=4, output_fields=4, patch_size=16, embed_dim=384, num_heads=6, processor_blocks=12)
    kw = dict(cfg, time_window=5, drop_path=0.2)
    m = get_model("filmavit", num_fluid_params=9, **kw) if film else get_model("avit", **kw)
    shapes = param_shapes(num_fluid_params=9 if film else None, **cfg)
    sd = m.state_dict()
    assert list(sd.keys()) == list(shapes.keys())
    assert all(tuple(sd[k].shape) == tuple(v) for k, v in shapes.items())
    assert sum(p.numel() for p in m.parameters()) == (28906602 if film else 28898904)
    assert len(list(m.buffers())) == 0
    # layer-scale / frequency-scalar / attention-scale initialisation (upstream attention.py:30,54-56,142,183-193)
    assert float(sd["blocks.0.temporal.gamma"][0]) == pytest.approx(1e-6)
    assert float(sd["blocks.3.spatial.low_freq_scalar"].abs().max()) == 0.0
    assert float(sd["blocks.3.spatial.attn_scale_factor_x"].min()) == 1.0


def test_modules_refuse_cpu_tensors():
    from bubbleformer_b200 import get_model
    m = get_model("filmavit", input_fields=4, output_fields=4, time_window=5, patch_size=16, embed_dim=128,
                  num_heads=2, processor_blocks=1, num_fluid_params=9)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 5, 4, 64, 64), torch.zeros(1, 9))
