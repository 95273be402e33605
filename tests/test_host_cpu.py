"""CPU-only tests: C-ABI surface, data-contract layer, scenario generator, sharding math."""
import os
import re
import sys

import numpy as np
import pytest

import helpers
from conftest import GOLDEN, ROOT
from harmonic_power_flow_b200 import _lib, netio, scenarios
from harmonic_power_flow_b200 import dist as hdist


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "hpf_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(hpf_[a-z_A-Z0-9]+)\s*\(", hdr))
    assert declared, "header parse failed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()                        # binds every symbol or raises
    for name in declared:
        assert hasattr(lib, name)
    m = re.search(r"#define HPF_ABI_VERSION (\d+)", hdr)
    assert lib.hpf_abi_version() == int(m.group(1)) == _lib.ABI_VERSION


def test_create_fails_loudly_without_gpu():
    import ctypes as C
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.hpf_create(C.byref(h), 0)
    assert rc == _lib.HPF_E_CUDA and not h.value
    assert b"CUDA" in lib.hpf_last_error(None)
    from harmonic_power_flow_b200 import BatchSolver
    net, _, _ = None, None, None
    with pytest.raises(RuntimeError):
        BatchSolver(helpers.packed_from_files("net3", 5, True, _tmp())[0])


def _tmp():
    import tempfile
    return tempfile.mkdtemp(prefix="hpf_test_")


def test_null_handle_is_rejected():
    lib = _lib.load()
    assert lib.hpf_dim_N(None) == 0
    assert lib.hpf_jacobian_stride(None) == 0
    assert lib.hpf_destroy(None) == 0
    assert lib.hpf_build_Y(None, None, None) == _lib.HPF_E_INVALID


@pytest.mark.parametrize("case", ["net3_c_h25", "net2_c_h19", "net2ev_c_h19", "net2_uc_h51",
                                  "net1_c_h25", "net2ev_uc_h19"])
def test_loader_matches_reference_per_unit_data(case, tmp_path):
    d = helpers.load_case(case)
    net, st, _ = helpers.packed_from_files(str(d["net"]), int(d["h_max"]), bool(d["coupled"]), tmp_path,
                                           julia_schema=str(d["net"]) == "net1")
    assert (net.n, net.m, net.c) == (int(d["n"]), int(d["m"]), int(d["c"]))
    assert np.array_equal(net.harmonics, d["harmonics"])
    assert np.array_equal(net.P, d["P"]) and np.array_equal(net.Q, d["Q"])
    # the NE CSV text round-trips bit-exactly and the p.u. conversion equals HG:301-308
    assert np.array_equal(net.I_N, d["I_N"])
    YN = net.Y_N[net.dev_of_nl_bus]
    assert np.array_equal(YN, d["Y_N"])
    assert net.N == 2 * net.n * net.H - 1 - net.c


def test_ne_lookup_is_case_insensitive_and_checks_harmonics(tmp_path):
    helpers.write_ne_csvs(str(tmp_path))
    assert netio.find_ne_file("SMPS", str(tmp_path)).endswith("smps_NE.csv")
    with pytest.raises(FileNotFoundError):
        netio.find_ne_file("toaster", str(tmp_path))
    pb, pl = helpers.write_net_csvs("net3", str(tmp_path))
    st = netio.Settings(H_MAX=101, ne_dir=str(tmp_path))          # file only has orders <= 99
    buses, lines, m, n, c = netio.init_network(pb, pl, st)
    with pytest.raises(KeyError):
        netio.import_Norton_Equivalents(buses, True, st)


def test_bus_order_is_checked(tmp_path):
    pb, pl = helpers.write_net_csvs("net3", str(tmp_path))
    txt = open(pb).read().splitlines()
    txt[2], txt[3] = txt[3], txt[2]          # PQ before PV
    open(pb, "w").write("\n".join(txt))
    with pytest.raises(ValueError):
        netio.init_network(pb, pl, netio.Settings())


def test_trailing_blank_line_and_vnom_column(tmp_path):
    pb, pl = helpers.write_net_csvs("net3", str(tmp_path))
    open(pl, "a").write("\n")                 # HPF/net2_lines.csv ends with a blank line
    rows = open(pb).read().splitlines()
    rows = [r + (";V_nom" if i == 0 else ";230") for i, r in enumerate(rows)]
    open(pb, "w").write("\n".join(rows))
    buses, lines, m, n, c = netio.init_network(pb, pl, netio.Settings())
    assert (m, n, c) == (3, 4, 2) and len(lines) == 4


@pytest.mark.parametrize("name", ["net3_c_h25_tight", "net3_c_h25_wide", "net2ev_c_h19_tight"])
def test_scenario_generator_reproduces_golden_inputs(name, tmp_path):
    d = helpers.load_set(name)
    net, _, _ = helpers.packed_from_files(str(d["net"]), int(d["h_max"]), bool(d["coupled"]), tmp_path)
    S = len(d["seed"])
    spread = "tight" if name.endswith("tight") else "wide"
    P, Q, I_N = scenarios.make_batch(net, S, spread)
    assert np.array_equal(P.T, d["P"]) and np.array_equal(Q.T, d["Q"])
    assert np.abs(np.moveaxis(I_N, 2, 0) - d["I_N"]).max() < 1e-15
    # pure function of (B, spread, seed0), also past the per-seed prefix
    a = scenarios.make_batch(net, 40, spread)
    b = scenarios.make_batch(net, 40, spread)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert np.array_equal(a[0][:, :8], P[:, :8])


def test_shard_bounds_cover_batch_exactly():
    for B in (0, 1, 7, 64, 65536, 65537):
        for w in (1, 2, 3, 4, 8):
            spans = [hdist.shard_bounds(B, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
            per = hdist.shard_size(B, w)
            assert all(hi - lo <= per for lo, hi in spans)
            for s in range(0, B, max(1, B // 13)):
                r = s // per
                assert spans[r][0] <= s < spans[r][1]


def test_binary_ne_format_round_trip(tmp_path):
    """<device>_NE.npz (next-4: the faster NE format the reference leaves as a TODO, HG:282) gives
    exactly the table the CSV gives, and import_Norton_Equivalents picks it up when no CSV exists."""
    import numpy as np
    from conftest import GOLDEN
    import os
    from harmonic_power_flow_b200 import netio
    dev = np.load(os.path.join(GOLDEN, "ne_devices.npz"))
    args = (dev["smps__freqs"], dev["smps__Y_N_c"], dev["smps__I_N_c"], dev["smps__Y_N_uc"], dev["smps__I_N_uc"])
    d_csv, d_npz = tmp_path / "csv", tmp_path / "npz"
    d_csv.mkdir(); d_npz.mkdir()
    netio.write_ne_csv(str(d_csv / "smps_NE.csv"), *args)
    netio.write_ne_npz(str(d_npz / "SMPS_NE.npz"), *args)
    a, b = netio.read_ne_csv(str(d_csv / "smps_NE.csv")), netio.read_ne_npz(str(d_npz / "SMPS_NE.npz"))
    assert list(a.index) == list(b.index) and list(a.columns) == list(b.columns)
    assert np.array_equal(a.to_numpy(), b.to_numpy())
    import pandas as pd
    buses = pd.DataFrame({"type": ["slack", "nonlinear"], "component": ["generator", "smps"]})
    for coupled in (True, False):
        st1, st2 = netio.Settings(H_MAX=25, ne_dir=str(d_csv)), netio.Settings(H_MAX=25, ne_dir=str(d_npz))
        n1 = netio.import_Norton_Equivalents(buses, coupled, st1)["smps"]
        n2 = netio.import_Norton_Equivalents(buses, coupled, st2)["smps"]
        assert np.array_equal(np.asarray(n1[0]), np.asarray(n2[0])) and np.array_equal(np.asarray(n1[1]), np.asarray(n2[1]))


def test_bus_auto_sorting(tmp_path):
    """next-4: a bus file in arbitrary order (the reference's TODO, HG:114) is brought into
    slack, PV.., PQ.., nonlinear.. order with consistent line end points; Y(h) of the sorted
    network is the permuted Y(h) of a hand-sorted one (checked with the oracle's assembly)."""
    import numpy as np
    import pandas as pd
    import hpf_oracle as O
    from harmonic_power_flow_b200 import netio
    st = netio.Settings(H_MAX=7)
    buses = pd.DataFrame({"ID": [1, 2, 3, 4, 5], "type": ["PQ", "nonlinear", "slack", "PV", "PQ"],
                          "component": ["l1", "smps", "gen", "g2", "l2"], "S": [0.0] * 5,
                          "P": [0.1, 0.25, 0.0, -0.2, 0.05], "Q": [0.1, 0.1, 0.0, 0.0, 0.02],
                          "X_sh": [0.0, 0.0, 3e-5, 0.0, 0.0]})
    lines = pd.DataFrame({"ID": [1, 2, 3, 4, 5], "fromID": [3, 1, 4, 5, 2], "toID": [1, 4, 5, 2, 3],
                          "R": [0.003, 0.006, 0.003, 0.004, 0.005], "X": [0.003, 0.025, 0.006, 0.01, 0.02],
                          "G": [0.0] * 5, "B": [0.0] * 5})
    with pytest.raises(ValueError):
        netio.check_bus_order(buses, 1, 2)
    b2, l2, order = netio.sort_network(buses, lines)
    assert list(b2["type"]) == ["slack", "PV", "PQ", "PQ", "nonlinear"] and list(order) == [2, 3, 0, 4, 1]
    assert list(b2["ID"]) == [1, 2, 3, 4, 5] and list(b2["component"]) == ["gen", "g2", "l1", "l2", "smps"]
    netio.check_bus_order(b2, 4, 2)
    # same physical branches: end points map through the permutation
    pos = {int(old): k + 1 for k, old in enumerate(buses["ID"].to_numpy()[order])}
    assert list(l2["fromID"]) == [pos[v] for v in lines["fromID"]] and list(l2["toID"]) == [pos[v] for v in lines["toID"]]
    mk = lambda b, l: O.Net(n=5, m=4, c=2, harmonics=np.array([1, 3, 5, 7]), line_from=l["fromID"].to_numpy(),
                            line_to=l["toID"].to_numpy(), R=l["R"].to_numpy(), X=l["X"].to_numpy(),
                            G=l["G"].to_numpy(), B=l["B"].to_numpy(), X_sh=b["X_sh"].to_numpy(),
                            P=b["P"].to_numpy(), Q=b["Q"].to_numpy())
    Y_sorted = O.build_admittance_matrices(mk(b2, l2))
    Y_orig = O.build_admittance_matrices(mk(buses, lines))
    assert np.abs(Y_sorted - Y_orig[:, order][:, :, order]).max() <= 1e-12 * np.abs(Y_orig).max()


def test_committed_bench_lines_keep_the_driver_contract():
    """The bench lines kept under profiles/ (GPU arm and reference arm of the final binary) carry every key
    of the driver contract; the GPU arm's traffic figure was captured from the sources it was run with."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ours = json.loads(open(os.path.join(root, "profiles", "r2_bench_e_1gpu.json")).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in ours, k
    assert ours["warmup"] >= 3 and ours["dtype"] == "f64" and ours["gpu_launches"] > 0
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(ours["e2e"])
    assert ours["e2e"]["h2d_bytes_per_step"] > 0 and ours["e2e"]["d2h_bytes_per_step"] > 0
    assert ours["e2e"]["value"] < ours["value"]
    rf = ours["roofline"]
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(rf)
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert rf["traffic_provenance"]["stale"] is False
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(ours["cpu_baseline"])
    assert all(ours["parity"]["gpu_within_floor"].values())
    ref = json.loads(open(os.path.join(root, "profiles", "r2_bench_reference_arm.json")).read().strip().splitlines()[-1])
    assert ref["impl"] == "reference" and ref["metric"] == ours["metric"] and ref["unit"] == ours["unit"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["cpu_baseline"]["kind"] in ("port", "reference")


def test_committed_ncu_summary_was_taken_from_the_current_sources():
    """profiles/r2_ncu_kernels.csv (the source of roofline.traffic) records the hash of the CUDA sources it was
    captured from: it must be the hash of the sources in this tree (bench.py marks the figure stale otherwise)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    first = open(bench.NCU_CSV).readline()
    assert first.startswith("#") and ("src_hash=" + bench.source_hash()) in first, first
    traffic, prov = bench.ncu_traffic("harm_hw_kernel")
    assert prov["stale"] is False and traffic and 5e7 < traffic < 5e8
