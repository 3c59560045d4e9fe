"""The C-ABI library loads and exports every symbol include/cmr_b200.h declares (no compute calls:
this runs without a GPU), and the ctypes signature table mirrors the header one to one."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cmr_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"CMR_API\s+[\w\s\*]+?\b(cmr_\w+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    for must in ("cmr_observe", "cmr_episode_prepare", "cmr_step", "cmr_reward", "cmr_to_disentangled",
                 "cmr_farthest_point_sample", "cmr_knn", "cmr_query_ball_point", "cmr_index_points",
                 "cmr_square_distance", "cmr_group_points", "cmr_cloud_mean"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from cmr_agent_b200 import _lib, build
    build.build_library()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in cmr_b200.h but not exported"
    assert lib.cmr_abi_version() == 5


def test_ctypes_table_matches_header():
    from cmr_agent_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    src = open(HEADER).read()
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", src, re.S)
        assert m, name
        params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(args), f"{name}: header has {len(params)} parameters, ctypes table {len(args)}"


def test_pure_host_entry_points_without_a_gpu():
    from cmr_agent_b200 import _lib
    lib = _lib.load()
    assert lib.cmr_workspace_bytes(0, 1, 1, 1) == 0
    n = lib.cmr_workspace_bytes(32, 40960, 64, 5120)
    assert n >= 32 * 40960 * 64 * 4 and n % 256 == 0
    assert lib.cmr_reward_scratch_bytes(8) > 0
    assert b"align" in lib.cmr_error_string(-2)
    # argument validation happens before any CUDA call
    assert lib.cmr_observe(None, None, None, None, None, None, None, 1, 1, 64, 1, 1, None, None, None, None, None) == -1
    assert lib.cmr_knn(None, None, 1, 1, 1, 1, None, None) == -1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from cmr_agent_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.CmrError):
        _lib.load()


def test_cpu_tensors_are_rejected_not_computed():
    import torch
    from cmr_agent_b200 import _lib, environment, pointnet_util, synth
    data = synth.make_batch(1, num_pt=64, img_h=16, img_w=16)
    with pytest.raises(_lib.CmrError):
        environment.observation_from_a_pose(data, torch.eye(4).repeat(1, 1, 1))
    with pytest.raises(_lib.CmrError):
        pointnet_util.farthest_point_sample(torch.zeros(1, 8, 3), 2)
    with pytest.raises(_lib.CmrError):
        pointnet_util.square_distance(torch.zeros(1, 8, 3), torch.zeros(1, 8, 3))


def test_graft_entry_build_runs_on_a_cpu_box():
    """The driver's "does it build" check: compiles (or finds) the library and the oracle, checks the ABI version."""
    import __graft_entry__ as g
    g.build()
