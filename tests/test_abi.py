"""CPU: the C-ABI shared library loads without a GPU and exports every symbol include/ia_b200.h declares; the
ctypes signature table matches the header; host-only entry points work (no compute calls without a GPU)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def header_functions():
    text = open(os.path.join(ROOT, "include", "ia_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = re.findall(r"\b(?:int32_t|int64_t|const char \*)\s*\*?\s*(ia_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S)
    out = {}
    for name, args in decls:
        args = args.strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        out[name] = n
    return out


@pytest.fixture(scope="module")
def lib():
    from instant_angelo_b200 import _lib, build
    build.build()
    return _lib.load()


def test_header_symbols_exported_and_bound(lib):
    from instant_angelo_b200 import _lib
    funcs = header_functions()
    assert len(funcs) >= 25
    for name, n_args in funcs.items():
        assert hasattr(lib, name), f"{name} is declared in ia_b200.h but not exported by libia_b200.so"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
        assert len(_lib.SIGNATURES[name][1]) == n_args, f"{name}: header has {n_args} parameters, binding has {len(_lib.SIGNATURES[name][1])}"
    assert set(_lib.SIGNATURES) == set(funcs), set(_lib.SIGNATURES) ^ set(funcs)


def test_host_only_entry_points(lib):
    from instant_angelo_b200 import _lib
    assert lib.ia_abi_version() == 1
    assert isinstance(lib.ia_last_error_string(), bytes)
    plan = _lib.GridPlan()
    assert lib.ia_hashgrid_plan(16, 2, 19, 32, C.c_float(1.3195079107728942), C.byref(plan)) == 0
    # tcnn geometry: resolutions, dense/hash switch, offsets, totals
    assert list(plan.res)[:16] == [32, 43, 56, 74, 98, 129, 169, 223, 295, 389, 513, 676, 892, 1177, 1553, 2049]
    assert list(plan.hashed)[:16] == [0, 0, 0, 0] + [1] * 12
    assert list(plan.offset)[:5] == [0, 32768, 112280, 287896, 693120]
    assert plan.n_entries == 6984576 and plan.n_params == 13969152
    assert lib.ia_hashgrid_plan(16, 2, 21, 32, C.c_float(1.3195079107728942), C.byref(plan)) == 0
    assert plan.n_params == 49405968 and list(plan.hashed)[:6] == [0, 0, 0, 0, 0, 1]        # L5 knife-edge: res 129 => hashed
    # invalid arguments: status code + message, never an exception
    assert lib.ia_hashgrid_plan(99, 2, 19, 32, C.c_float(1.3), C.byref(plan)) == -1
    assert b"n_levels" in lib.ia_last_error_string()
    desc = _lib.MlpDesc(3, 2.0, -1.0, 32, 2, 64, 65, _lib.IA_ACT_SOFTPLUS100, _lib.IA_ACT_NONE, _lib.IA_MLP_FP32)
    assert lib.ia_mlp_param_count(C.byref(desc)) == 64 * 35 + 64 + 64 * 64 + 64 + 65 * 64 + 65
    assert lib.ia_occ_workspace_bytes(128 ** 3) >= 4 * 128 ** 3


def test_plan_matches_oracle(lib):
    from instant_angelo_b200 import ops
    from oracle import tcnn_ref as tc
    for cfg in [dict(n_levels=16, n_features=2, log2_hashmap_size=19, base_resolution=32, per_level_scale=1.3195079107728942),
                dict(n_levels=16, n_features=2, log2_hashmap_size=21, base_resolution=32, per_level_scale=1.3195079107728942),
                dict(n_levels=8, n_features=2, log2_hashmap_size=12, base_resolution=4, per_level_scale=1.5),
                dict(n_levels=12, n_features=2, log2_hashmap_size=15, base_resolution=16, per_level_scale=2.0)]:
        ref = tc.grid_plan(**cfg)
        got = ops.make_grid_plan(**cfg)
        L = cfg["n_levels"]
        assert list(got.res)[:L] == ref.res and list(got.size)[:L] == ref.size
        assert list(got.offset)[:L + 1] == ref.offset and [bool(h) for h in list(got.hashed)[:L]] == ref.hashed
        assert [float(s) for s in list(got.scale)[:L]] == ref.scale


def test_product_refuses_cpu_tensors():
    """No CPU fallback: the drop-in modules raise on non-CUDA inputs like nerfacc / tcnn do."""
    import torch
    from instant_angelo_b200 import make, nerfacc_api
    from instant_angelo_b200.config import to_config
    from tests.golden.scenes import golden_model_config
    model = make("neus", to_config(golden_model_config()))
    model.train()
    model.background_color = torch.ones(3)
    with pytest.raises(NotImplementedError, match="cuda"):
        model(torch.rand(4, 6))
    with pytest.raises(NotImplementedError, match="cuda"):
        nerfacc_api.ray_aabb_intersect(torch.rand(4, 3), torch.rand(4, 3), torch.tensor([-1.0, -1, -1, 1, 1, 1]))
    with pytest.raises(NotImplementedError, match="cuda"):
        nerfacc_api.render_weight_from_alpha(torch.rand(4, 1), ray_indices=torch.zeros(4, dtype=torch.int32), n_rays=1)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under instant_angelo_b200/ may import it."""
    pkg = os.path.join(ROOT, "instant_angelo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports the oracle"
                assert "from oracle" not in text and "import oracle" not in text, f
