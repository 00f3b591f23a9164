"""CPU tests of the product's host logic and of the C-ABI surface (no compute calls without a GPU)."""
import ctypes
import dataclasses
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from ssrs_b200 import build, _native
    lib_path = build.build()
    header = open(os.path.join(ROOT, "include", "ssrs_b200.h")).read()
    declared = set(re.findall(r"SSRS_API\s+[\w\s\*]+?\b(ssrs_\w+)\s*\(", header))
    assert declared and declared == set(_native.EXPORTS), declared ^ set(_native.EXPORTS)
    lib = ctypes.CDLL(lib_path)
    for name in declared:
        assert hasattr(lib, name), name
    lib.ssrs_abi_version.restype = ctypes.c_int
    assert lib.ssrs_abi_version() == 1
    _native.load()


def test_bindings_match_the_header_prototypes():
    """Every ctypes prototype in ssrs_b200/_native.py against the C prototype in include/ssrs_b200.h: same number of
    parameters, same kind per position (pointer / int / int64 / uint64 / float / double) and the same return type — an
    argument dropped or widened in a binding otherwise shows up as a garbage stream handle at run time."""
    from ssrs_b200 import _native
    header = open(os.path.join(ROOT, "include", "ssrs_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    protos = dict((m.group(2), (m.group(1).strip(), m.group(3)))
                  for m in re.finditer(r"SSRS_API\s+([\w\s\*]+?)\b(ssrs_\w+)\s*\(([^;]*?)\)\s*;", header, flags=re.S))
    assert set(protos) == set(_native.EXPORTS)

    def kind_of_c(decl):
        decl = " ".join(decl.split())
        if "*" in decl:
            return "ptr"
        base = decl.rsplit(" ", 1)[0] if " " in decl else decl          # drop the parameter name
        base = base.replace("const ", "").replace("unsigned long long", "uint64_t").strip()
        return {"int": "int", "int32_t": "int", "int64_t": "i64", "long long": "i64", "uint64_t": "u64", "float": "f32",
                "double": "f64"}[base]

    def kind_of_ctypes(t):
        if t in (ctypes.c_void_p, ctypes.c_char_p) or (isinstance(t, type) and issubclass(t, ctypes._Pointer)):
            return "ptr"
        return {ctypes.c_int: "int", ctypes.c_int32: "int", ctypes.c_int64: "i64", ctypes.c_uint64: "u64",
                ctypes.c_float: "f32", ctypes.c_double: "f64"}[t]

    for name, (ret, params) in sorted(protos.items()):
        restype, argtypes = _native.EXPORTS[name]
        c_params = [] if params.strip() in ("", "void") else [p for p in params.split(",")]
        assert len(c_params) == len(argtypes), (name, len(c_params), len(argtypes))
        for i, (c, t) in enumerate(zip(c_params, argtypes)):
            assert kind_of_c(c) == kind_of_ctypes(t), (name, i, c.strip(), t)
        assert kind_of_c(ret + " x") == kind_of_ctypes(restype), (name, ret, restype)


def test_integration_sketch_calls_match_the_header():
    """The ctypes sketch in INTEGRATION.md (what a maintainer of the reference would paste): it parses, and every
    `lib.ssrs_*(...)` call passes as many arguments as the header's prototype has parameters."""
    import ast
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    assert blocks
    header = re.sub(r"/\*.*?\*/", " ", open(os.path.join(ROOT, "include", "ssrs_b200.h")).read(), flags=re.S)
    nparams = {m.group(1): (0 if m.group(2).strip() in ("", "void") else m.group(2).count(",") + 1)
               for m in re.finditer(r"SSRS_API\s+[\w\s\*]+?\b(ssrs_\w+)\s*\(([^;]*?)\)\s*;", header, flags=re.S)}
    seen = set()
    for block in blocks:
        tree = ast.parse(block)
        for node in ast.walk(tree):
            if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr.startswith("ssrs_") \
                    and isinstance(node.func.value, ast.Name) and node.func.value.id == "lib":
                n = 0
                for arg in node.args:
                    # `*map(_p, out)` spreads the four output rasters of ssrs_updraft
                    n += 4 if isinstance(arg, ast.Starred) else 1
                assert node.func.attr in nparams, node.func.attr
                assert n == nparams[node.func.attr], (node.func.attr, n, nparams[node.func.attr])
                seen.add(node.func.attr)
    assert {"ssrs_updraft", "ssrs_potential_solve", "ssrs_step_tracks", "ssrs_interleave_fields"} <= seen


def test_simulator_constructor_host_part(tmp_path, monkeypatch):
    """The part of `Simulator.__init__` that runs before the first CUDA call (reference simulator.py:40-127): output
    directories, the JSON dump of the configuration, grid size and bounds, and the attribute surface plot code relies on
    (SURVEY App. E).  The constructor is stopped at its CUDA gate — there is no CPU path behind it."""
    import json
    from datetime import datetime
    import ssrs_b200.simulator as S
    from ssrs_b200 import Config, Simulator

    class Gate(Exception):
        pass

    def gate():
        raise Gate()
    monkeypatch.setattr(S.N, "require_cuda", gate)
    z = np.zeros((120, 160), np.float32)
    for mode, cases, dtimes in (("uniform", None, None),
                                ("snapshot", {"y2014m12d01h15": (8.0, 270.0)}, [datetime(2014, 12, 1, 15)]),
                                ("seasonal", {"case00": (8.0, 270.0), "case01": (6.0, 200.0)}, None)):
        cfg = Config(run_name=f"host_{mode}", out_dir=str(tmp_path), sim_mode=mode, region_width_km=(16., 12.), resolution=100.)
        sim = Simulator.__new__(Simulator)
        with pytest.raises(Gate):
            sim.__init__(cfg, elevation=z, wind_cases=cases)
        assert sim.gridsize == (120, 160) and sim.bounds == (0.0, 0.0, 15900.0, 11900.0)
        assert sim.extent == (0.0, 15900.0, 0.0, 11900.0)
        assert os.path.isdir(sim.mode_data_dir) and sim.mode_data_dir.endswith(os.path.join("data", mode))
        assert os.path.isdir(sim.mode_fig_dir)
        with open(os.path.join(str(tmp_path), cfg.run_name, f"{cfg.run_name}.json")) as f:
            dumped = json.load(f)
        assert dumped["sim_mode"] == mode and dumped["resolution"] == 100.0 and "wtk_layers" not in dumped
        assert sim.wtk_layers["wspeed"] == "windspeed_100m" and sim.wtk_layers["temperature"] == "temperature_100m"
        assert sim.wtk is None and sim.dtimes == dtimes
        assert Simulator.lonlat_crs == "EPSG:4326" and Simulator.time_format == "y%Ym%md%dh%H"
    with pytest.raises(ValueError):            # terrain is injected, never downloaded
        Simulator(Config(run_name="noelev", out_dir=str(tmp_path)))
    with pytest.raises(ValueError):
        sim.get_wtk_locs()
    sim._wind_points = (np.arange(3.0), np.arange(3.0) + 1, None)
    assert np.array_equal(sim.get_wtk_locs()[1], [1.0, 2.0, 3.0])


def test_host_only_entry_points_answer_without_a_gpu():
    """Entry points that are pure host arithmetic: halo transport query of a missing communicator, phase count of the
    phased stepper, workspace and table sizes."""
    from ssrs_b200 import _native
    lib = _native.load()
    assert lib.ssrs_comm_halo_mode(None) == -1
    assert lib.ssrs_step_phase_count(5000, 6000, 0) == 34 and lib.ssrs_step_phase_count(3, 3, 0) == 0
    assert lib.ssrs_walk_workspace_bytes(1000) == 2 * 1000 * 16 + 4096
    assert lib.ssrs_walk_table_bytes(5000, 6000) == 5000 * 6000 * 64


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ssrs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "/root/reference" not in src.replace("`/root/reference", "").replace("(`/root/reference", "") \
                    or f.endswith(".py"), f


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ssrs_b200 import layers, _native
    with pytest.raises(_native.NativeError):
        layers.compute_slope_degrees(np.zeros((8, 8), np.float32), 10.0)


def test_config_contract():
    from ssrs_b200 import Config
    c = Config()
    names = [f.name for f in dataclasses.fields(c)]
    assert len(names) == 34 and names[0] == "run_name" and names[-1] == "fig_dpi"
    assert c.updraft_threshold == 0.75 and c.track_dirn_restrict == 1 and c.track_start_region == (5, 55, 1, 2)
    assert c.turbine_mrkr_styles[0] == "1k"
    s = str(c)
    assert ":::: Simulating tracks\nmovement_model = fluidflow" in s       # the reference's positional section break
    c2 = dataclasses.replace(c, track_count=5, sim_mode="snapshot")
    assert c2.track_count == 5 and c.track_count == 1000


def test_host_helpers_against_golden(golden):
    from ssrs_b200 import movmodel as mm
    g = golden("tables")
    for i in range(9):
        assert np.array_equal(mm.get_track_restrictions(i // 3 - 1, i % 3 - 1), g["masks"][i])
    for th, w in zip(g["thetas"], g["dirw"]):
        assert np.array_equal(mm.get_directional_probs(th * np.pi / 180.0), w)
    assert np.array_equal(mm.neighbour_delta_norms_inv, g["norms_inv"])
    assert np.array_equal(np.array(mm.neighbour_deltas), g["deltas"])
    for th in (0, 30, 45, 90, 180, 270, 315, -45):
        for shp in ((23, 31), (60, 50)):
            bn, be = mm.MovModel(th, shp).get_boundary_nodes()
            assert np.array_equal(bn, g[f"bn_{th}_{shp[0]}x{shp[1]}"]) and np.array_equal(be, g[f"be_{th}_{shp[0]}x{shp[1]}"])
    np.random.seed(4)
    r, c = mm.get_starting_indices(64, (5, 55, 1, 2), "random", (60., 50.), 100.)
    assert np.array_equal(r, g["start_random_rows"]) and np.array_equal(c, g["start_random_cols"])
    r, c = mm.get_starting_indices(37, (5, 55, 1, 2), "structured", (60., 50.), 100.)
    assert np.array_equal(r, g["start_struct_rows"]) and np.array_equal(c, g["start_struct_cols"])
    with pytest.raises(ValueError):
        mm.get_starting_indices(10, (60, 5, 1, 2), "random", (60., 50.), 100.)
    with pytest.raises(ValueError):
        mm.get_starting_indices(10, (5, 55, 1, 2), "spiral", (60., 50.), 100.)
    assert mm.move_away_from_boundary(0, 0, 50, 60) == (2, 2)
    assert mm.move_away_from_boundary(1, 1, 50, 60) == (3, 1)          # row <= 1 but col <= 0: asymmetric
    assert mm.move_away_from_boundary(48, 58, 50, 60) == (46, 56)
    assert mm.harmonic_mean(0.0, 3.0, 1e-8) == 1e-8 and mm.harmonic_mean(1.0, 3.0) == 1.5


def test_assemble_matches_oracle_operator():
    """MovModel.assemble_sparse_linear_system (API parity) encodes the same links as the oracle's operator,
    including the last-column S/SW factor swap."""
    from oracle import oracle_np as O
    from ssrs_b200 import movmodel as mm
    nrow, ncol = 9, 7
    ri, ci, fa = mm.MovModel(0.0, (nrow, ncol)).assemble_sparse_linear_system()
    assert ri.dtype == np.uint32 and fa.dtype == np.float32
    K = np.ones((nrow, ncol))
    g = O.edge_weights(K)
    dense = np.zeros((nrow * ncol, nrow * ncol))
    dense[ri, ci] = 1.0 / fa.astype(np.float64)
    for d in range(9):
        dr, dc = d // 3 - 1, d % 3 - 1
        for r in range(nrow):
            for c in range(ncol):
                if d == 4 or not (0 <= r + dr < nrow and 0 <= c + dc < ncol):
                    continue
                i, j = c * nrow + r, (c + dc) * nrow + (r + dr)
                assert dense[i, j] == g[d, r, c], (r, c, dr, dc)


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py's stdout contract: ONE JSON line, whatever libraries write to file descriptor 1 (it is re-pointed at
    stderr for the run).  The reference arm runs on CPU; a small grid keeps it to a few seconds."""
    import json
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--rows", "300", "--cols", "400", "--resolution", "100", "--tracks-per-gpu", "2000"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:2000]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "track-steps/sec" and d["value"] > 0
    # the reference's own modules when they are reachable (/root/reference here, the staged oracle/_ref on the GPU box)
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["e2e"]["h2d_bytes_per_step"] == 0


def test_phase_schedule_and_walk_policy():
    """Host arithmetic behind the phased launches: the schedule is geometric (x1.25) from one crossing of the short side up
    to max_moves; the walk policy depends on the global track count and the grid only."""
    from ssrs_b200 import _native, movmodel as mm
    lib = _native.load()
    assert lib.ssrs_step_phase_count(5000, 6000, 0) == 34          # 5000, 6252, ... < 7.5e6, + the open-ended last phase
    assert lib.ssrs_step_phase_count(120, 160, 0) == 8             # first cut at 1024 steps, max_moves = 4800
    assert lib.ssrs_step_phase_count(120, 160, 64) > 15
    assert lib.ssrs_step_phase_count(3, 3, 0) == 0
    assert mm.walk_pays_off(100_000, (5000, 6000)) and not mm.walk_pays_off(1000, (500, 600))
    assert not mm.walk_pays_off(10 ** 6, (5000, 6000), memory_parameter=2)
    assert lib.ssrs_walk_table_bytes(5000, 6000) == 5000 * 6000 * 64
    assert lib.ssrs_walk_workspace_bytes(1000) >= 2 * 1000 * 16


def test_reference_staging(tmp_path, monkeypatch):
    """oracle/ref_loader.stage_reference copies the reference's two hot-path modules byte for byte (git-ignored
    oracle/_ref/), and the loader finds them there when /root/reference is absent (the GPU box)."""
    from oracle import ref_loader as R
    if not os.path.isfile("/root/reference/ssrs/movmodel.py"):
        pytest.skip("reference tree not present")
    monkeypatch.setattr(R, "STAGED_ROOT", str(tmp_path / "_ref"))
    assert R.stage_reference()
    for m in R.HOT_PATH_MODULES:
        assert open(tmp_path / "_ref" / "ssrs" / m, "rb").read() == open(f"/root/reference/ssrs/{m}", "rb").read()
    gi = open(os.path.join(ROOT, ".gitignore")).read()
    assert "oracle/_ref/" in gi
